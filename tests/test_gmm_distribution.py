"""Distribution test of the vectorised GMM generator (mdcommunity_b200/gmm.py) against statistics recorded from the
REFERENCE generator (GMM.py + Hyperbolic.py; fixture tests/golden/gmm_stats.npz, made by make_golden_gmm_stats.py).
The two generators consume their random streams in different orders, so graphs are equal in distribution only: every
recorded per-graph statistic is compared with a two-sample Kolmogorov-Smirnov test, the pooled degree distribution
bin by bin (max absolute difference of the normalised histograms < 0.02).  CPU only."""
import os

import numpy as np
import pytest
from scipy import stats

from conftest import GOLDEN
from mdcommunity_b200 import gmm

Z = np.load(os.path.join(GOLDEN, "gmm_stats.npz"))
FAMILIES = sorted(k[4:] for k in Z.files if k.startswith("fam_"))
P_MIN = 1e-3  # per test; 8 families x 9 tests: a correct generator fails one of them with probability < 0.1


def _ours(name, n, count):
    g, gamma, kbar = [float(x) for x in Z["par_" + name]]
    rows, hist = [], np.zeros(n, np.int64)
    for i in range(count):
        rng = np.random.default_rng(7000 + i)
        kb = (rng.uniform(2.0, 10.0), rng.uniform(2.0, 10.0)) if kbar < 0 else (kbar, kbar)
        e = gmm.sample_layers(n, rng, g=g, gamma=(gamma, gamma), kbar=kb)
        deg = [np.bincount(x.ravel(), minlength=n) for x in e]
        key = [set(map(tuple, np.sort(x, 1).tolist())) for x in e]
        cc = np.corrcoef(deg[0], deg[1])[0, 1] if deg[0].std() > 0 and deg[1].std() > 0 else 0.0
        rows.append([len(e[0]), len(e[1]), deg[0].max(), deg[1].max(), (deg[0] == 0).sum(), (deg[1] == 0).sum(),
                     len(key[0] & key[1]), cc])
        for d in deg:
            hist += np.bincount(d, minlength=n)[:n]
    return np.array(rows, dtype=np.float64), hist


@pytest.mark.parametrize("name", FAMILIES)
def test_generator_matches_reference_statistics(name):
    n = int(Z["N"])
    ref, ref_hist = Z["fam_" + name], Z["hist_" + name]
    ours, hist = _ours(name, n, 120)
    cols = [str(c) for c in Z["columns"]]
    for j, c in enumerate(cols):
        p = stats.ks_2samp(ref[:, j], ours[:, j]).pvalue
        assert p > P_MIN, (name, c, p, ref[:, j].mean(), ours[:, j].mean())
    # pooled degree distribution: bins 0..14 and a tail bin
    def binned(h):
        return np.concatenate([h[:15], [h[15:].sum()]]).astype(np.float64)
    a, b = binned(ref_hist), binned(hist)
    # the pooled histogram mixes 60 / 120 graphs whose degrees are correlated through kappa, so nodes are not independent
    # draws (a chi-square test would overstate the sample size): compare the normalised distributions directly
    pa, pb = a / a.sum(), b / b.sum()
    assert np.abs(pa - pb).max() < 0.02, (name, np.abs(pa - pb).max())
