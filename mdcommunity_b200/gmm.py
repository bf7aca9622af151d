"""Vectorised geometric multiplex model (GMM) generator for synthetic two-layer graphs.

Same model as the reference's generator (/root/reference/code/MultiDismantler_unit_cost/GMM.py:6-68 and
Hyperbolic.py:18-117: S1 hidden-variable model per layer, inter-layer correlation of hidden degrees (nu) and of
angular coordinates (g)), restated with numpy array operations: the reference's O(N^2) Python pair loop
(Hyperbolic.py:101-117, 7.6 s at N=1000) becomes one vectorised upper-triangle draw.  The random stream is
consumed in a different order, so graphs are equal in distribution, not edge-for-edge.
"""
import numpy as np
from scipy.special import erf, erfinv, lambertw


def _kmin(kbar, gamma):
    return kbar * (gamma - 2.0) / (gamma - 1.0)


def sample_layers(N, rng, g=0.5, nu=0.2, gamma=(2.5, 2.5), kbar=(6.0, 6.0), T=(0.4, 0.4)):
    """Returns two undirected edge arrays [E_l, 2] (i < j)."""
    kmin1, kmin2 = _kmin(kbar[0], gamma[0]), _kmin(kbar[1], gamma[1])
    g1, g2 = gamma
    kappa1 = kmin1 * np.power(1.0 - rng.random(N), 1.0 / (1.0 - g1))
    if nu == 1:
        kappa2 = kmin2 * np.power(kappa1 / kmin1, (1.0 - g1) / (1.0 - g2))
    elif nu == 0:
        kappa2 = kmin2 * np.power(1.0 - rng.random(N), 1.0 / (1.0 - g2))
    else:
        phi = -np.log(1.0 - np.power(kmin1 / kappa1, g1 - 1.0))
        z = 1.0 / kmin1 * np.power(phi, nu / (nu - 1.0)) * np.power(kappa1, -g1)
        z = z * (kmin1 * np.power(kappa1, g1) - np.power(kmin1, g1) * kappa1)
        zr = z * rng.random(N)
        zr = (nu / (1.0 - nu)) * lambertw(np.power(zr, (nu - 1.0) / nu) / (nu / (1.0 - nu)))
        zr = np.power(zr, 1.0 / (1.0 - nu)) - np.power(phi, 1.0 / (1.0 - nu))
        zr = np.exp(-np.power(zr, 1.0 - nu))
        kappa2 = np.real(kmin2 * np.power(1.0 - zr, 1.0 / (1.0 - g2)))
    theta1 = 2.0 * np.pi * rng.random(N)
    if g == 1:
        theta2 = theta1.copy()
    elif g == 0:
        theta2 = 2.0 * np.pi * rng.random(N)
    else:
        sigma0 = min(N / (4.0 * np.pi), 100.0)
        sigma = sigma0 * (1.0 / g - 1.0)
        l = np.sqrt(2.0) * sigma * erfinv((-1.0 + 2.0 * rng.random(N)) * erf(N / (2 * np.sqrt(2) * sigma)))
        theta2 = (theta1 + 2 * np.pi * l / N) % (2 * np.pi)
    out = []
    iu, ju = np.triu_indices(N, k=1)
    for kappa, theta, Tl, kb in ((kappa1, theta1, T[0], kbar[0]), (kappa2, theta2, T[1], kbar[1])):
        dtheta = N / (2 * np.pi) * np.abs(np.pi - np.abs(np.pi - np.abs(theta[iu] - theta[ju])))
        mu = np.sin(Tl * np.pi) / (2 * np.pi * Tl * kb)
        r = dtheta / (mu * kappa[iu] * kappa[ju])
        p = 1.0 / (1.0 + np.power(r, 1.0 / Tl))
        keep = rng.random(iu.size) < p
        out.append(np.stack([iu[keep], ju[keep]], 1).astype(np.int64))
    return out


def gmm_graph(N, seed, cost="unit", **kw):
    """Graph(N) of the reference: unit cost draws kbar_l ~ U(2,10) (U/GMM.py:17,23), degree cost fixes kbar = 6
    (D/GMM.py:16,20); g = 0.5, nu = 0.2, gamma = 2.5, T = 0.4 in both."""
    from .graph import Graph_test
    rng = np.random.default_rng(seed)
    if "kbar" not in kw:
        kw["kbar"] = (6.0, 6.0) if cost == "degree" else (rng.uniform(2.0, 10.0), rng.uniform(2.0, 10.0))
    e1, e2 = sample_layers(N, rng, **kw)
    return Graph_test(e1, e2, num_nodes=N, cost=cost)


def random_sparse_layers(N, n_layers, n_edges, seed=0):
    """Config C4: independent Erdos-Renyi-like layers with exactly n_edges distinct undirected edges each."""
    out = []
    for l in range(n_layers):
        rng = np.random.default_rng(seed * 131 + l)
        have = np.zeros(0, np.int64)
        while have.size < n_edges:
            need = int((n_edges - have.size) * 1.1) + 16
            u = rng.integers(0, N, need, dtype=np.int64)
            v = rng.integers(0, N, need, dtype=np.int64)
            ok = u != v
            a, b = np.minimum(u[ok], v[ok]), np.maximum(u[ok], v[ok])
            have = np.unique(np.concatenate([have, a * N + b]))
        have = rng.permutation(have)[:n_edges]
        out.append(np.stack([have // N, have % N], 1))
    return out
