#!/usr/bin/env python
"""Statistics of the REFERENCE's synthetic-graph generator (GMM.py + Hyperbolic.py), recorded as a fixture for the
distribution test of mdcommunity_b200/gmm.py (tests/test_gmm_distribution.py).

Runs the unmodified reference generator (read-only at /root/reference; matplotlib stubbed because Hyperbolic.py:4-6
imports it).  Cannot run on the GPU box; its output tests/golden/gmm_stats.npz is committed.

    python tests/golden/make_golden_gmm_stats.py

Per family (the generator knobs of U/GMM.py:10-25 -- g, gamma, kbar; nu = 0.2, T = 0.4 fixed) and per graph:
edge counts of both layers, maximum degrees, number of nodes of degree 0, the inter-layer edge overlap |E1 & E2|,
the Pearson correlation of the two layers' degree sequences, and the pooled degree histogram of the family.
Families: the unit-cost default (kbar_l ~ U(2, 10), g = 0.5, gamma = 2.5: U/GMM.py), the degree-cost default
(kbar = 6: D/GMM.py) and the three testSynthetic sweeps (U/testSynthetic.py:20: data_g / data_gamma / data_k).
"""
import os
import random
import sys
import types

import numpy as np

REF = "/root/reference/code/MultiDismantler_unit_cost"
OUT = os.path.dirname(os.path.abspath(__file__))
N = 200
PER_FAMILY = 60

FAMILIES = {
    # name: (g, gamma, kbar or None for U(2,10))
    "unit_default": (0.5, 2.5, None),
    "degree_default": (0.5, 2.5, 6.0),
    "g0": (0.0, 2.5, 6.0), "g1": (1.0, 2.5, 6.0),
    "gamma2.1": (0.5, 2.1, 6.0), "gamma3.0": (0.5, 3.0, 6.0),
    "k4": (0.5, 2.5, 4.0), "k10": (0.5, 2.5, 10.0),
}


def install_shims():
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def reference_gmm(hyp, n, g, gamma, kbar):
    """The body of GMM() (U/GMM.py:6-68) with its parameter block exposed -- the calls into Hyperbolic.py are the
    reference's own, in the reference's order, so the random streams are consumed exactly as GMM() does."""
    nu, T = 0.2, 0.4
    kbar1 = random.uniform(2.0, 10.0) if kbar is None else kbar
    kbar2 = random.uniform(2.0, 10.0) if kbar is None else kbar
    kmin1 = hyp.CalculateKmin(kbar1, gamma)
    kmin2 = hyp.CalculateKmin(kbar2, gamma)
    kappa1 = hyp.SampleKappa(n, kmin1, gamma)
    kappa2 = hyp.SampleConditionalKappa(n, nu, kappa1[:], kmin1, gamma, kmin2, gamma)
    theta1 = hyp.SampleTheta(n)
    theta2 = hyp.SampleConditionalTheta(n, g, theta1[:])
    links1 = hyp.CreateNetworks(kappa1, theta1, T, kbar1)
    links2 = hyp.CreateNetworks(kappa2, theta2, T, kbar2)
    return links1, links2


def stats(n, l1, l2):
    e = [np.asarray(l, dtype=np.int64).reshape(-1, 2) for l in (l1, l2)]
    deg = [np.bincount(x.ravel(), minlength=n) for x in e]
    key = [set(map(tuple, np.sort(x, 1).tolist())) for x in e]
    cc = np.corrcoef(deg[0], deg[1])[0, 1] if deg[0].std() > 0 and deg[1].std() > 0 else 0.0
    return [len(e[0]), len(e[1]), deg[0].max(), deg[1].max(), int((deg[0] == 0).sum()), int((deg[1] == 0).sum()),
            len(key[0] & key[1]), cc], deg


def main():
    install_shims()
    sys.path.insert(0, REF)
    os.chdir(REF)
    import Hyperbolic as hyp
    out = {"N": np.int64(N), "columns": np.array(["E1", "E2", "dmax1", "dmax2", "iso1", "iso2", "overlap", "deg_corr"])}
    for name, (g, gamma, kbar) in FAMILIES.items():
        rows = []
        hist = np.zeros(N, np.int64)
        for i in range(PER_FAMILY):
            random.seed(1000 + i)
            np.random.seed(1000 + i)
            l1, l2 = reference_gmm(hyp, N, g, gamma, kbar)
            r, deg = stats(N, l1, l2)
            rows.append(r)
            for d in deg:
                hist += np.bincount(d, minlength=N)[:N]
        out["fam_" + name] = np.array(rows, dtype=np.float64)
        out["hist_" + name] = hist
        out["par_" + name] = np.array([g, gamma, -1.0 if kbar is None else kbar])
        print(name, np.mean(np.array(rows), 0).round(2))
    np.savez_compressed(os.path.join(OUT, "gmm_stats.npz"), **out)


if __name__ == "__main__":
    main()
