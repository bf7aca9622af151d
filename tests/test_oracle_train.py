"""Oracle restatement of one DQN training step (train_forward + calc_loss + backward) vs the golden vectors recorded
from the reference (tests/golden/make_golden_train.py -> tests/golden/train_unit.npz).  CPU only."""
import os

import numpy as np
import torch

from conftest import GOLDEN, load_golden_graph


def replay_states(z):
    from oracle import md_oracle as orc
    envs = []
    for case, t in zip(z["cases"], z["t"]):
        zc, _, g = load_golden_graph(os.path.join(GOLDEN, str(case) + ".npz"))
        env = orc.OracleEnv("unit")
        env.s0(g)
        for a in zc["greedy_actions"][:int(t)]:
            env.stepWithoutReward(int(a))
        envs.append(env)
    return envs


def test_train_loss_and_gradients_match_reference():
    from oracle import md_oracle as orc
    z = np.load(os.path.join(GOLDEN, "train_unit.npz"))
    w = orc.load_weights(os.path.join(GOLDEN, "weights_unit.npz"))
    w = {k: v.clone().requires_grad_(True) for k, v in w.items() if k != "last_w"}
    envs = replay_states(z)
    loss, q_pred, embs = orc.train_loss(w, envs, z["actions"], z["targets"])
    assert np.allclose(q_pred.detach().numpy(), z["q_pred"], rtol=2e-5, atol=1e-6)
    assert np.allclose(embs[0].detach().numpy(), z["emb0"], rtol=0, atol=2e-6)
    assert np.allclose(embs[1].detach().numpy(), z["emb1"], rtol=0, atol=2e-6)
    assert abs(float(loss) - float(z["loss"])) <= 1e-6 * abs(float(z["loss"])) + 1e-9
    loss.backward()
    for k in z.files:
        if not k.startswith("grad/"):
            continue
        g_ref = z[k]
        g = w[k[5:]].grad.numpy()
        scale = np.abs(g_ref).max() + 1e-12
        assert np.abs(g - g_ref).max() <= 2e-4 * scale + 1e-9, (k, np.abs(g - g_ref).max(), scale)
    # one Adam step (torch.optim.Adam defaults, lr = LEARNING_RATE) reproduces the reference's updated parameters
    opt = torch.optim.Adam(list(w.values()), lr=orc.LEARNING_RATE)
    opt.step()
    for k in z.files:
        if k.startswith("after/"):
            assert np.allclose(w[k[6:]].detach().numpy(), z[k], rtol=0, atol=2e-6), k


def test_nstep_returns():
    from oracle import md_oracle as orc
    r = [-0.5, -0.25, -0.125, -0.0625]
    out = orc.nstep_returns(r, 2)
    assert out[0] == (r[0] + r[1], 2, False) and out[1] == (r[1] + r[2], 3, False)
    assert out[2] == (r[2] + r[3], None, True) and out[3] == (r[3], None, True)
