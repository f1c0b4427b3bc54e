"""The REINFORCE oracle (oracle/reinforce_oracle.py) against vectors produced by the unmodified reference
(recnn/nn/models.py:76-184, recnn/nn/update/reinforce.py:10-65)."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import reinforce_oracle as RO
from tests import _reinforce as RG


@pytest.mark.parametrize("path", RG.FILES, ids=RG.IDS)
def test_select_action_quantities(path):
    c = RG.load(path)
    g = c["g"]
    for t in range(c["T"]):
        probs, _ = RO.discrete_forward(c["params"], g["states"][t])
        np.testing.assert_allclose(probs, g["probs"][t], rtol=2e-5, atol=1e-9)
        lp = RO.categorical_log_prob(g["probs"][t], c["pi_action"][t])
        np.testing.assert_allclose(lp, g["saved_log_probs"][t], rtol=1e-5, atol=1e-6)
        if c["method"] != RO.BASIC:
            blp = RO.categorical_log_prob(g["beta_probs"][t], c["beta_action"][t])
            np.testing.assert_allclose(RO.correction(lp, blp), g["correction"][t], rtol=2e-5)
        if c["method"] == RO.TOPK:
            np.testing.assert_allclose(RO.lambda_k(lp, c["K"]), g["lambda_k"][t], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("path", RG.FILES, ids=RG.IDS)
def test_returns(path):
    c = RG.load(path)
    np.testing.assert_allclose(RO.normalised_returns(c["g"]["rewards"]), c["g"]["returns"], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("path", RG.FILES, ids=RG.IDS)
def test_loss_gradient_and_sgd_step(path):
    c = RG.load(path)
    g = c["g"]
    ret = g["returns"][c["rows_step"]]
    loss, grads, _ = RO.reinforce_policy_grad(c["params"], c["rows_state"], c["rows_action"], c["rows_beta_logp"], ret,
                                              c["method"], c["K"])
    assert loss == pytest.approx(float(g["loss"]), rel=2e-5, abs=1e-5)
    names = {"w1": "linear1.weight", "b1": "linear1.bias", "w2": "linear2.weight", "b2": "linear2.bias"}
    for k, n in names.items():
        ref = g["grad." + n]
        scale = np.abs(ref).max()
        np.testing.assert_allclose(grads[k], ref, rtol=2e-4, atol=2e-6 * scale, err_msg=n)
        after = c["params"][k].astype(np.float64) - float(g["lr"]) * grads[k]
        np.testing.assert_allclose(after, g["after." + n], rtol=1e-5, atol=1e-6 * max(1.0, scale), err_msg=n)


def test_inverse_cdf_sampling_is_a_categorical_draw():
    rng = np.random.default_rng(5)
    probs = rng.dirichlet(np.ones(7), size=1).astype(np.float32)
    u = rng.random(20000).astype(np.float32)
    act, lp, margin = RO.categorical_sample(np.repeat(probs, u.size, 0), u)
    freq = np.bincount(act, minlength=7) / u.size
    np.testing.assert_allclose(freq, probs[0], atol=0.012)
    np.testing.assert_allclose(lp, np.log(probs[0].astype(np.float64) / probs[0].astype(np.float64).sum())[act], rtol=1e-6)
    assert (margin >= 0).all()
    # zero-probability outcomes are never drawn, u = 0 picks the first possible one
    p = np.asarray([[0.0, 0.0, 0.5, 0.0, 0.5]], np.float32)
    assert RO.categorical_sample(p, np.asarray([0.0], np.float32))[0][0] == 2
    assert RO.categorical_sample(p, np.asarray([0.5], np.float32))[0][0] == 4
    assert RO.categorical_sample(p, np.asarray([0.999], np.float32))[0][0] == 4


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("method", [RO.BASIC, RO.CORRECTED, RO.TOPK])
def test_vocabulary_sharded_formulation_equals_unsharded(world, method):
    """Design study for the 1M-item configuration: sharding linear2 / the softmax over ranks with an exchange of
    (max, sum, drawn logit) per row and an all-reduce of dh reproduces the unsharded loss and gradients."""
    rng = np.random.default_rng(world * 10 + method)
    S, H, I, R = 20, 24, 101, 37                      # 101 items: uneven shards (and an empty one at world 8? no: 13 each)
    p = RO.make_discrete_actor(rng, S, I, H)
    state = rng.normal(0, 1, (R, S)).astype(np.float32)
    action = rng.integers(0, I, R)
    action[:3] = [0, I - 1, I // 2]
    blp = np.log(rng.uniform(0.005, 0.02, R)).astype(np.float32)
    ret = rng.normal(0, 1, R).astype(np.float32)
    want_loss, want, _ = RO.reinforce_policy_grad(p, state, action, blp, ret, method, 7)
    shards = RO.shard_policy(p, world)
    got_loss, got = RO.sharded_policy_grad(shards, state, action, blp, ret, method, 7)
    assert got_loss == pytest.approx(want_loss, rel=1e-12, abs=1e-12)
    w2 = np.concatenate([g["w2"] for g in got], 0)
    b2 = np.concatenate([g["b2"] for g in got], 0)
    np.testing.assert_allclose(w2, want["w2"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(b2, want["b2"], rtol=1e-9, atol=1e-12)
    for g in got:
        np.testing.assert_allclose(g["w1"], want["w1"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(g["b1"], want["b1"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("path", RG.FILES, ids=RG.IDS)
def test_host_side_returns_of_the_mirror_match_the_reference(path, monkeypatch):
    """recnn_b200.nn.ChooseREINFORCE.__call__ forms the normalised discounted returns on the host (reinforce.py:44-52)
    before the one device call; with that call stubbed out the host arithmetic is checkable on a CPU box: the returns it
    hands over are the reference's, the optimizer is stepped once and the policy's lists are cleared."""
    import torch
    import recnn_b200
    from recnn_b200.nn.update import reinforce as mirror
    c = RG.load(path)
    seen = {}

    def fake_policy_loss(policy, returns, method):
        seen["returns"] = torch.as_tensor(returns).clone()
        seen["method"] = method
        return torch.tensor(1.5)

    monkeypatch.setattr(mirror, "_policy_loss", fake_policy_loss)
    policy = recnn_b200.nn.DiscreteActor(c["S"], c["I"], c["H"])
    policy.rewards = [torch.tensor(float(r)) for r in c["g"]["rewards"]]
    policy.saved_log_probs = [torch.zeros(1)] * c["T"]

    class Opt:
        steps = 0

        def step(self):
            Opt.steps += 1

    chooser = recnn_b200.nn.ChooseREINFORCE(getattr(recnn_b200.nn.ChooseREINFORCE, c["method_name"]))
    out = chooser(policy, Opt(), learn=True)
    assert float(out) == 1.5 and Opt.steps == 1 and seen["method"] == c["method"]
    np.testing.assert_allclose(seen["returns"].numpy(), c["g"]["returns"], rtol=1e-6, atol=1e-6)
    assert policy.rewards == [] and policy.saved_log_probs == [] and policy._saved == []
    with pytest.raises(TypeError):
        recnn_b200.nn.ChooseREINFORCE(lambda p, r: 0)(policy, Opt())
