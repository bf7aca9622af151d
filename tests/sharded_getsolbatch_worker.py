"""torchrun worker of tests/test_gpu_mirrors.py::test_sharded_getsolbatch_two_gpus (one rank per GPU, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mdcommunity_b200 as mdb
    from mdcommunity_b200 import gmm
    graphs = [gmm.gmm_graph(150 + 10 * i, 40 + i, cost="unit") for i in range(9)]
    dqn = mdb.MultiDismantler(cost="unit", device="cuda:%d" % local)
    dqn.LoadModel(os.path.join(ROOT, "tests", "golden", "weights_unit.npz"))
    score, sols, res = dqn.GetSolBatch(graphs)
    np.savez(os.path.join(out, "rank%d.npz" % dist.get_rank()), score=score, n_steps=res["n_steps"], actions=res["actions"],
             max_rank=res["max_rank"])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
