"""Config C5: DQN training step on a replay batch of 64 transitions drawn from epsilon = 1 random play on synthetic
two-layer graphs of N = 200 (unit cost): train_forward + calc_loss + backward + gradient all-reduce + Adam.
Run alone (1 GPU) or under torchrun (data-parallel, every rank takes the round-robin shard of the SAME batch).
Prints one JSON line on rank 0: training steps / s and transitions / s."""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdcommunity_b200 import gmm, train  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--graphs", type=int, default=16)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
graphs = [gmm.gmm_graph(args.n, 1000 + i, cost="unit") for i in range(args.graphs)]
tr = train.DQNTrainer("unit", seed=0)
tr.LoadModel(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_unit.npz"))))
mem = train.NStepReplayMem(100000)
sim = train.Simulator(tr, graphs, rng=random.Random(0))
sim.play(1.0, mem)                       # the same replay memory on every rank (same seeds)
rng = random.Random(1)
samples = [mem.sampling(args.batch, rng) for _ in range(args.steps + args.warmup)]
targets = [train.fit_targets(tr, s) for s in samples]
losses = []
for i, (s, t) in enumerate(zip(samples, targets)):
    if i == args.warmup:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
    losses.append(tr.fit(s.g_list, s.list_st, s.list_at, t))
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": "C5: DQN training step, replay batch %d, N=%d, unit cost" % (args.batch, args.n), "n_gpus": world,
                      "steps": args.steps, "train_steps_per_s": args.steps / dt, "transitions_per_s": args.steps * args.batch / dt,
                      "ms_per_step": 1e3 * dt / args.steps, "replay_entries": mem.count,
                      "first_loss": losses[0], "last_loss": losses[-1]}))
if world > 1:
    dist.destroy_process_group()
