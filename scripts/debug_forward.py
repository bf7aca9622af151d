"""Stage-by-stage comparison of the CUDA forward against the oracle on one golden graph (GPU box diagnostic)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C  # noqa: E402

import mdcommunity_b200 as mdb  # noqa: E402
from mdcommunity_b200 import _lib  # noqa: E402
from oracle import md_oracle as orc  # noqa: E402


def main(variant="unit", name="n100_s0", steps=3):
    path = os.path.join(ROOT, "tests", "golden", "%s_%s.npz" % (variant, name))
    z = np.load(path)
    g = mdb.Graph_test(z["edges0"], z["edges1"], num_nodes=int(z["num_nodes"]), cost=variant)
    batch = mdb.GraphBatch([g], cost=variant)
    wd = dict(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % variant)))
    w = mdb.QWeights(wd, batch.device)
    ro = mdb.BatchRollout(batch, w)
    ow = orc.load_weights(os.path.join(ROOT, "tests", "golden", "weights_%s.npz" % variant))
    og = orc.MultiplexGraph(int(z["num_nodes"]), [z["edges0"], z["edges1"]],
                            weights=[z["weights0"], z["weights1"]] if variant == "degree" else None)
    env = orc.OracleEnv(variant)
    env.s0(og)
    print("max_rank gpu", g.max_rank, "oracle", og.max_rank, "golden", int(z["max_rank"]))
    lib = ro.lib
    for t in range(steps):
        tr = {}
        q, act_ids, off = orc.forward_q(ow, [env], variant, trace=tr)
        ids = act_ids[0]
        gl = ro.label.cpu().numpy()[:g.num_nodes]
        ol = np.where(env.removed, -1, env.labels)
        print("step", t, "labels equal", np.array_equal(gl, ol), "n_act", int(ro.n_act[0]), len(ids))
        gact = ro.act_list.cpu().numpy()[:int(ro.n_act[0])]
        print("  active equal", np.array_equal(gact, ids))
        for l in range(2):
            print("  deg equal L%d" % l, np.array_equal(ro.deg[l].cpu().numpy()[:g.num_nodes], env.alive_degree(l)))
        gp, sp, wp, kp, st = ro._g(), C.byref(ro.s), ro.w.ptr(), C.byref(ro.k), ro._stream()
        _lib.check(lib.md_embed_init(gp, sp, wp, kp, st), "init")
        for l in range(2):
            d = ro.h[0][l].cpu().numpy()[ids] - tr["h"][(l, 0)]
            print("  H0 L%d maxdiff %.3e" % (l, np.abs(d).max()))
        for it in range(3):
            _lib.check(lib.md_embed_iter(gp, sp, wp, kp, it, st), "iter")
            for l in range(2):
                d = ro.h[(it + 1) & 1][l].cpu().numpy()[ids] - tr["h"][(l, it + 1)]
                ps = ro.psum.view(3, 2, -1, 64)[it, l].cpu().numpy()
                nt = (len(ids) + 127) // 128
                print("  H%d L%d maxdiff %.3e   ypool maxdiff %.3e" % (it + 1, l, np.abs(d).max(),
                      np.abs(ps[:nt].sum(0) - tr["ypool"][(l, it)][0]).max()))
        _lib.check(lib.md_graph_embed(gp, sp, wp, kp, st), "gembed")
        gc = ro.gconst.cpu().numpy()[0]
        n = tr["n"]
        yf = [tr["fused"][l][n:] for l in range(2)]
        s_or = [float(yf[l][0] @ ow["cross_product"].numpy()[:, 0]) for l in range(2)]
        print("  s gpu", gc[0:2], "oracle", s_or)
        print("  gate gpu", gc[4:6], "oracle", tr["gate_softmax"][0])
        _lib.check(lib.md_qhead(gp, sp, wp, kp, st), "qhead")
        gi, gq = ro.q_values(0)
        print("  q maxdiff %.3e (max|q| %.3e) argmax gpu %d oracle %d golden %d" % (
            np.abs(gq - q).max(), np.abs(q).max(), gi[np.argmax(gq)], ids[np.argmax(q)], int(z["greedy_actions"][t])))
        a = int(z["greedy_actions"][t])
        acts = torch.tensor([[a]], dtype=torch.int32, device=batch.device)
        ro.step(acts, 0)
        env.stepWithoutReward(a)
        print("  after step: rank gpu", int(ro.rank[0]), "oracle", env.rank_list[-1], "score gpu %.17g oracle %.17g" % (
            float(ro.score[0]), env.score))


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["unit"]), *(sys.argv[2:3] or ["n100_s0"]))
