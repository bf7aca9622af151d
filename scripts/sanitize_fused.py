"""compute-sanitizer target: a few forwards + steps of md_embed_fused on a mixed batch (large and small graphs), packed and
legacy.  Usage: compute-sanitizer --tool memcheck python scripts/sanitize_fused.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import _lib, gmm  # noqa: E402

cost = sys.argv[1] if len(sys.argv) > 1 else "unit"
sizes = [1000, 40, 60, 300, 25, 90, 130, 200, 33, 256, 257, 128, 64, 45]
graphs = [gmm.gmm_graph(n, 700 + i, cost=cost) for i, n in enumerate(sizes)]
batch = mdb.GraphBatch(graphs, "cuda:0", cost)
w = mdb.QWeights(dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % cost))), batch.device)
lib = _lib.lib()
for pack, flags in ((0, 0), (-1, 0), (0, 1 << 19)):
    lib.md_debug_set_fused_pack(pack)
    lib.md_debug_set_flags(flags)
    ro = mdb.BatchRollout(batch, w, fused=True)
    for t in range(6):
        ro.forward()
        ro.step()
    torch.cuda.synchronize()
    print("pack", pack, "flags", flags, "ok; n_act", ro.n_act.cpu().numpy().tolist())
