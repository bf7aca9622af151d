"""Where one DQN training step (config C5) spends its time: wall clock per phase of DQNTrainer.fit with a device
synchronize after each (1 GPU).  Prints one JSON line."""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdcommunity_b200 import dense, gmm, train  # noqa: E402
from mdcommunity_b200.graph import GraphBatch  # noqa: E402
from mdcommunity_b200.rollout import BatchRollout  # noqa: E402

torch.cuda.set_device(0)
graphs = [gmm.gmm_graph(200, 1000 + i, cost="unit") for i in range(16)]
tr = train.DQNTrainer("unit", seed=0)
tr.LoadModel(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_unit.npz"))))
mem = train.NStepReplayMem(100000)
train.Simulator(tr, graphs, rng=random.Random(0)).play(1.0, mem)
rng = random.Random(1)
acc = {}


def tick(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


STEPS, WARM = 20, 3
for i in range(STEPS + WARM):
    s = mem.sampling(64, rng)
    tgt_np = train.fit_targets(tr, s)
    if i == WARM:
        acc.clear()
    torch.cuda.synchronize()
    t = time.perf_counter()
    batch = GraphBatch(s.g_list, tr.device, "unit")
    t = tick("GraphBatch (host CSR of 64 graphs + H2D)", t)
    ro = BatchRollout(batch, None, fused=False)
    t = tick("BatchRollout (state buffers)", t)
    train.set_states(ro, s.list_st)
    t = tick("set_states (prefix replay on the device env, %d steps max)" % max(len(c) for c in s.list_st), t)
    tb = train.TrainBatch(ro)
    t = tick("TrainBatch (compact CSR operands, index plumbing)", t)
    o = tb.operands
    rows = tb.rows_of(s.list_at, batch.S)
    tgt = torch.as_tensor(tgt_np.reshape(-1)).float().to(tr.device)
    params = tr.net.tensors()
    c = dense.forward(params, o)
    t = tick("forward (hand-written kernels)", t)
    q_pred = c["q"].index_select(0, rows)
    loss, lc = dense.loss_forward(o, q_pred, c["F"], tgt, 64, tb.edge_num, train.ALPHA)
    dq_pred, dF = dense.loss_backward(o, lc)
    dq = torch.zeros(tb.n, dtype=torch.float32, device=tr.device)
    dq.index_copy_(0, rows, dq_pred)
    t = tick("loss forward + backward", t)
    tr.optimizer.zero_grad()
    dense.backward(params, o, c, dq, dF, tr.net.grad_tensors())
    t = tick("backward (hand-written kernels)", t)
    tr.optimizer.adopt_flat_grads()
    tr.optimizer.step()
    t = tick("Adam (one kernel)", t)
print(json.dumps({"config": "C5 phases, 1 GPU, batch 64, N=200", "rows_n": tb.n, "ms_per_step": {k: 1e3 * v / STEPS for k, v in acc.items()},
                  "total_ms": 1e3 * sum(acc.values()) / STEPS}))
