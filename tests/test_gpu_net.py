"""Model-level boundary: mdcommunity_b200.net.MultiDismantler_net with the reference's state_dict keys and
test_forward / train_forward signatures, fed with PrepareBatchGraph-style COO operands."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_cases, load_golden_graph
from test_oracle_train import replay_states

pytestmark = pytest.mark.gpu


def _operands(envs, dev):
    """What SetuppredAll / SetupTrain hand to the network, built from oracle env states."""
    from oracle import md_oracle as orc
    w = orc.load_weights(os.path.join(GOLDEN, "weights_unit.npz"))
    with torch.no_grad():
        q, act_ids, off, ex = orc.forward_torch(w, envs, "unit")
    n, B = ex["n"], len(envs)
    gid = torch.from_numpy(np.repeat(np.arange(B), np.diff(off))).long()
    n2n = [{"index": torch.stack([ex["row"][l], ex["col"][l]]).to(dev), "value": torch.ones(ex["row"][l].numel()).to(dev),
            "m": n, "n": n} for l in range(2)]
    subg = [{"index": torch.stack([gid, torch.arange(n)]).to(dev), "value": torch.ones(n).to(dev), "m": B, "n": n} for _ in range(2)]
    rep = [{"index": torch.stack([torch.arange(n), gid]).to(dev), "value": torch.ones(n).to(dev), "m": n, "n": B} for _ in range(2)]
    aux = torch.from_numpy(np.stack([e.status()[2] for e in envs])).to(dev)
    return q, act_ids, off, n2n, subg, rep, aux


def _net(dev):
    from mdcommunity_b200.net import MultiDismantler_net
    net = MultiDismantler_net(device=dev)
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, "weights_unit.npz")).items()}
    assert len(sd) == 14
    missing, unexpected = net.load_state_dict(sd, strict=True)  # the reference's 14-key checkpoint, unchanged
    assert not missing and not unexpected
    return net


def test_test_forward_matches_reference_q():
    from oracle import md_oracle as orc
    dev = torch.device("cuda:0")
    net = _net(dev)
    envs, zs = [], []
    for p in golden_cases("unit", 200):
        z, _, g = load_golden_graph(p)
        e = orc.OracleEnv("unit")
        e.s0(g)
        envs.append(e)
        zs.append(z)
    q_orc, act_ids, off, n2n, subg, rep, aux = _operands(envs, dev)
    with torch.no_grad():
        q = net.test_forward(None, subg, n2n, rep, aux, adj=None, v_adj=None)[:, 0].cpu().numpy()
    for b, z in enumerate(zs):  # the reference's own f32 Q at the first greedy step
        o = z["greedy_off"]
        qg = z["greedy_q"][o[0]:o[1]]
        assert np.array_equal(act_ids[b], z["greedy_active"][o[0]:o[1]])
        assert np.all(np.abs(q[off[b]:off[b + 1]] - qg) <= 1e-4 * np.abs(qg) + 1e-6)


def test_train_forward_and_backward_match_reference():
    dev = torch.device("cuda:0")
    net = _net(dev)
    z = np.load(os.path.join(GOLDEN, "train_unit.npz"))
    envs = replay_states(z)
    _, act_ids, off, n2n, subg, rep, aux = _operands(envs, dev)
    rows = [int(off[b] + np.nonzero(act_ids[b] == int(a))[0][0]) for b, a in enumerate(z["actions"])]
    B = len(envs)
    act_sel = [{"index": torch.tensor([list(range(B)), rows]).to(dev), "value": torch.ones(B).to(dev), "m": B, "n": n2n[0]["m"]}
               for _ in range(2)]
    q_pred, cur = net.train_forward(None, subg, n2n, act_sel, aux, adj=None, v_adj=None)
    assert np.allclose(q_pred.detach().cpu().numpy(), z["q_pred"], rtol=1e-4, atol=1e-6)
    assert np.allclose(cur[0].detach().cpu().numpy(), z["emb0"], rtol=0, atol=5e-6)
    assert np.allclose(cur[1].detach().cpu().numpy(), z["emb1"], rtol=0, atol=5e-6)
    # calc_loss (MultiDismantler_torch.py:410-431) written with the operator-level spmm, then backward
    from mdcommunity_b200 import ops
    tgt = torch.as_tensor(z["targets"]).float().to(dev)
    loss = torch.nn.functional.mse_loss(tgt, q_pred)
    for l in range(2):
        idx = n2n[l]["index"]
        n = n2n[l]["m"]
        lap_index = torch.cat([torch.stack([torch.arange(n, device=dev)] * 2), idx], 1)
        deg = torch.bincount(idx[0], minlength=n).float()
        lap_value = torch.cat([deg, -torch.ones(idx.shape[1], device=dev)])
        recons = 2 * torch.trace(torch.matmul(cur[l].t(), ops.spmm(lap_index, lap_value, n, n, cur[l])))
        loss = loss + float(z["alpha"]) * recons / n2n[l]["value"].sum()
    assert abs(float(loss.detach()) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"])) + 1e-9
    loss.backward()
    named = dict(net.named_parameters())
    for k in z.files:
        if k.startswith("grad/"):
            g_ref = z[k]
            g = named[k[5:]].grad.cpu().numpy().reshape(g_ref.shape)
            scale = np.abs(g_ref).max() + 1e-12
            assert np.abs(g - g_ref).max() <= 1e-3 * scale + 1e-9, (k, np.abs(g - g_ref).max(), scale)
