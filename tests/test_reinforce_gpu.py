"""REINFORCE policy side on the GPU (recnn_b200/csrc/reinforce.cuh through the public API) against the oracle and the
reference's own vectors (tests/golden/reinforce_*.npz).  Tolerances: fp32 with 3xTF32 contractions against a float64
oracle -- 2e-5 relative on probabilities / log-probs / losses, gradients 2e-4 of the tensor's largest entry."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import recnn_b200
from oracle import recnn_oracle as O
from oracle import reinforce_oracle as RO
from tests import _reinforce as RG

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = {"w1": "linear1.weight", "b1": "linear1.bias", "w2": "linear2.weight", "b2": "linear2.bias"}


def make_policy(p, S, H, I):
    m = recnn_b200.nn.DiscreteActor(S, I, H)
    with torch.no_grad():
        m.linear1.weight.copy_(torch.from_numpy(p["w1"]))
        m.linear1.bias.copy_(torch.from_numpy(p["b1"]))
        m.linear2.weight.copy_(torch.from_numpy(p["w2"]))
        m.linear2.bias.copy_(torch.from_numpy(p["b2"]))
    return m.to(DEV)


def grads_of(m):
    return {"w1": m.linear1.weight.grad, "b1": m.linear1.bias.grad, "w2": m.linear2.weight.grad, "b2": m.linear2.bias.grad}


@pytest.mark.parametrize("S,H,I,N", [(13, 16, 37, 6), (52, 64, 300, 10), (1290, 256, 5000, 96), (1290, 2048, 5000, 10)])
def test_forward_matches_oracle(S, H, I, N):
    rng = np.random.default_rng(S + I)
    p = RO.make_discrete_actor(rng, S, I, H)
    state = rng.normal(0, 1, (N, S)).astype(np.float32)
    m = make_policy(p, S, H, I)
    got = m(torch.from_numpy(state)).cpu().numpy()
    want, _ = RO.discrete_forward(p, state)
    np.testing.assert_allclose(got, want, rtol=3e-5, atol=1e-10)
    np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-5)


def test_sampling_replays_uniforms_like_the_oracle():
    rng = np.random.default_rng(11)
    n, items = 512, 5000
    probs = rng.dirichlet(np.full(items, 0.05), size=n).astype(np.float32)
    probs[3, :100] = 0.0                              # leading zero-probability outcomes are never drawn
    u = rng.random(n).astype(np.float32)
    u[0], u[1] = 0.0, 0.99999994
    m = recnn_b200.nn.DiscreteActor(4, items, 4).to(DEV)
    m.uniform_source = lambda k: torch.from_numpy(u)
    act, lp = m._sample(torch.from_numpy(probs).to(DEV))
    want_a, want_lp, margin = RO.categorical_sample(probs, u)
    sure = margin > 2e-6
    assert sure.mean() > 0.95
    got_a = act.cpu().numpy()
    assert np.array_equal(got_a[sure], want_a[sure])
    assert (probs[np.arange(n), got_a] > 0).all()
    np.testing.assert_allclose(lp.cpu().numpy()[sure], want_lp[sure], rtol=1e-5, atol=1e-6)
    # log_prob of given actions
    lp2 = m._log_prob(torch.from_numpy(probs).to(DEV), torch.from_numpy(want_a))
    np.testing.assert_allclose(lp2.cpu().numpy(), want_lp, rtol=1e-5, atol=1e-6)
    with pytest.raises(IndexError):
        m._log_prob(torch.from_numpy(probs).to(DEV), torch.full((n,), items, dtype=torch.int64))


def test_philox_draws_are_categorical_and_repeatable():
    torch.manual_seed(1234)
    probs = torch.tensor([[0.1, 0.0, 0.2, 0.3, 0.4]], device=DEV).repeat(40000, 1).contiguous()
    m = recnn_b200.nn.DiscreteActor(4, 5, 4).to(DEV)
    a1, lp1 = m._sample(probs)
    freq = torch.bincount(a1, minlength=5).float().cpu().numpy() / probs.shape[0]
    np.testing.assert_allclose(freq, [0.1, 0.0, 0.2, 0.3, 0.4], atol=0.01)
    np.testing.assert_allclose(lp1.cpu().numpy(), np.log(np.asarray([0.1, 1, 0.2, 0.3, 0.4]))[a1.cpu().numpy()], rtol=1e-5)
    a2, _ = m._sample(probs)
    assert not torch.equal(a1, a2)                    # the draw counter advances
    m2 = recnn_b200.nn.DiscreteActor(4, 5, 4).to(DEV)
    b1, _ = m2._sample(probs)
    assert torch.equal(a1, b1)                        # same seed, same draw index -> same draws


@pytest.mark.parametrize("path", RG.FILES, ids=RG.IDS)
def test_choose_reinforce_matches_reference_vectors(path):
    """ChooseREINFORCE(method)(policy, SGD, learn=True) on the reference's recorded draws: loss, every .grad and the
    stepped parameters against what the unmodified reference produced."""
    c = RG.load(path)
    g = c["g"]
    m = make_policy(c["params"], c["S"], c["H"], c["I"])
    for t in range(c["T"]):
        rec = {"state": torch.from_numpy(g["states"][t]).to(DEV), "action": torch.from_numpy(c["pi_action"][t]).to(DEV),
               "beta_log_prob": None, "K": c["K"]}
        if c["method"] != RO.BASIC:
            rec["beta_log_prob"] = torch.from_numpy(c["rows_beta_logp"].reshape(c["T"], c["N"])[t]).to(DEV)
        m._saved.append(rec)
        m.saved_log_probs.append(torch.zeros(c["N"], device=DEV))
        m.rewards.append(torch.tensor(g["rewards"][t], device=DEV))
    opt = torch.optim.SGD(m.parameters(), lr=float(g["lr"]))
    chooser = recnn_b200.nn.ChooseREINFORCE(getattr(recnn_b200.nn.ChooseREINFORCE, c["method_name"]))
    loss = chooser(m, opt, learn=True)
    assert torch.is_tensor(loss) and loss.dim() == 0
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-4, abs=1e-4)
    assert len(m.saved_log_probs) == 0 and len(m.rewards) == 0 and len(m._saved) == 0      # gc()
    got = grads_of(m)
    for k, name in NAMES.items():
        ref = g["grad." + name]
        scale = np.abs(ref).max()
        np.testing.assert_allclose(got[k].cpu().numpy(), ref, rtol=1e-3, atol=2e-4 * scale, err_msg=name)
        after = dict(m.named_parameters())[name].detach().cpu().numpy()
        np.testing.assert_allclose(after, g["after." + name], rtol=1e-4, atol=1e-5 * max(1.0, scale), err_msg=name)


@pytest.mark.parametrize("method", ["basic_reinforce", "reinforce_with_correction", "reinforce_with_TopK_correction"])
@pytest.mark.parametrize("S,H,I,N,T", [(1290, 256, 5000, 32, 4), (1290, 2048, 5000, 10, 10)])
def test_policy_gradient_tensor_core_sizes(method, S, H, I, N, T):
    """The notebook's shapes (1290 -> 2048 -> 5000 items, 10 rows x 10 env steps) and a wider batch: loss and gradients
    against the float64 oracle; built-in arena optimizer steps the same parameters."""
    rng = np.random.default_rng(T * 7 + H)
    p = RO.make_discrete_actor(rng, S, I, H)
    m = make_policy(p, S, H, I)
    states = rng.normal(0, 1, (T, N, S)).astype(np.float32)
    actions = rng.integers(0, I, (T, N))
    beta_lp = np.log(rng.uniform(1e-4, 5e-4, (T, N))).astype(np.float32)
    rewards = rng.normal(0, 1, T).astype(np.float32)
    mid = RO.METHODS[method]
    for t in range(T):
        m._saved.append({"state": torch.from_numpy(states[t]).to(DEV), "action": torch.from_numpy(actions[t]).to(DEV),
                         "beta_log_prob": None if mid == RO.BASIC else torch.from_numpy(beta_lp[t]).to(DEV), "K": 10})
        m.rewards.append(torch.tensor(rewards[t], device=DEV))
    opt = recnn_b200.optim.Adam(m.parameters(), lr=1e-3)
    before = m.linear2.weight.detach().clone()
    loss = recnn_b200.nn.ChooseREINFORCE(getattr(recnn_b200.nn.ChooseREINFORCE, method))(m, opt, learn=True)
    ret = RO.normalised_returns(rewards)[np.repeat(np.arange(T), N)]
    want_loss, want, _ = RO.reinforce_policy_grad(p, states.reshape(T * N, S), actions.reshape(-1),
                                                  None if mid == RO.BASIC else beta_lp.reshape(-1), ret, mid, 10)
    assert float(loss) == pytest.approx(want_loss, rel=2e-4, abs=1e-4 * (1 + abs(want_loss)))
    got = grads_of(m)
    for k in NAMES:
        scale = np.abs(want[k]).max()
        assert scale > 0
        err = np.abs(got[k].cpu().numpy() - want[k]).max()
        assert err <= 3e-4 * scale, (k, err, scale)
    assert opt.steps_taken() == 1
    assert not torch.equal(before, m.linear2.weight.detach())


def test_select_action_bookkeeping_matches_the_oracle():
    """_select_action_with_TopK_correction with replayed uniforms: the values appended to saved_log_probs / correction /
    lambda_k are the oracle's for the oracle's draws, with both action_source settings (models.py:137-140)."""
    rng = np.random.default_rng(3)
    S, H, I, N, K = 52, 64, 300, 16, 7
    p = RO.make_discrete_actor(rng, S, I, H)
    beta_w = rng.normal(0, 0.3, (I, S)).astype(np.float32)
    state = rng.normal(0, 1, (N, S)).astype(np.float32)
    for source in ({"pi": "pi", "beta": "beta"}, {"pi": "beta", "beta": "beta"}, {"pi": "pi", "beta": "pi"}):
        m = make_policy(p, S, H, I)
        m.action_source = dict(source)
        us = [rng.random(N).astype(np.float32) for _ in range(2)]
        it = iter(us)
        m.uniform_source = lambda n: torch.from_numpy(next(it))
        bw = torch.from_numpy(beta_w).to(DEV)
        beta = lambda s, action=None: torch.softmax(s @ bw.T, dim=1)       # noqa: E731  (the caller's behaviour policy)
        probs = m._select_action_with_TopK_correction(torch.from_numpy(state).to(DEV), beta, None, K,
                                                      recnn_b200.utils.DummyWriter(), 0)
        pi_probs, _ = RO.discrete_forward(p, state)
        z = state.astype(np.float64) @ beta_w.astype(np.float64).T
        e = np.exp(z - z.max(1, keepdims=True))
        beta_probs = e / e.sum(1, keepdims=True)
        np.testing.assert_allclose(probs.cpu().numpy(), pi_probs, rtol=3e-5)
        a_pi, _, m1 = RO.categorical_sample(pi_probs, us[0])
        a_beta, _, m2 = RO.categorical_sample(beta_probs, us[1])
        if min(m1.min(), m2.min()) < 2e-6:
            continue
        pi_action = a_pi if source["pi"] == "pi" else a_beta
        beta_action = a_beta if source["beta"] == "beta" else a_pi
        lp = RO.categorical_log_prob(pi_probs, pi_action)
        blp = RO.categorical_log_prob(beta_probs, beta_action)
        np.testing.assert_allclose(m.saved_log_probs[0].cpu().numpy(), lp, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(m.correction[0].cpu().numpy(), RO.correction(lp, blp), rtol=2e-4)
        np.testing.assert_allclose(m.lambda_k[0].cpu().numpy(), RO.lambda_k(lp, K), rtol=2e-4)
        assert np.array_equal(m._saved[0]["action"].cpu().numpy(), pi_action)
        np.testing.assert_allclose(m._saved[0]["beta_log_prob"].cpu().numpy(), blp, rtol=1e-4, atol=1e-6)


def _dense_batch(rng, N, S, I):
    action = rng.integers(0, I, N)
    one_hot = np.zeros((N, I), np.float32)
    one_hot[np.arange(N), action] = 1
    return {"state": rng.normal(0, 1, (N, S)).astype(np.float32), "action": one_hot,
            "reward": rng.integers(1, 6, N).astype(np.float32) - 3, "next_state": rng.normal(0, 1, (N, S)).astype(np.float32),
            "done": (rng.random(N) < 0.1).astype(np.float32)}


def test_critic_step_against_discrete_target_policy():
    """value_update with a DiscreteActor target policy (reinforce.py:92-102): loss and the stepped critic against the
    oracle; eval-mode critic (no dropout) so that no masks are involved."""
    rng = np.random.default_rng(21)
    S, H, I, N = 52, 64, 300, 24
    pp = RO.make_discrete_actor(rng, S, I, H)
    cp = O.make_critic(rng, S, I, H, 0.3)
    policy = make_policy(pp, S, H, I)
    from tests._cuda import load_net, dump_net
    value = load_net(recnn_b200.nn.Critic(S, I, H, 0.3), cp, DEV).eval()
    algo = recnn_b200.nn.Reinforce(policy, value)
    algo.nets["value_net"].eval()
    algo.optimizers["value_optimizer"] = recnn_b200.optim.Adam(value.parameters(), lr=1e-3)
    o_nets = {"value_net": O.copy_net(cp), "target_value_net": O.copy_net(cp), "target_policy_net": pp}
    o_opts = {"value_optimizer": O.make_optimizer("adam", lr=1e-3)}
    params = dict(algo.params)
    for step in range(3):
        b = _dense_batch(rng, N, S, I)
        want, _ = RO.value_update(b, params, o_nets, o_opts, None, learn=True)
        got = recnn_b200.nn.value_update({k: torch.from_numpy(v) for k, v in b.items()}, params, algo.nets, algo.optimizers,
                                         torch.device(DEV), {}, learn=True, step=step)
        assert float(got) == pytest.approx(float(want), rel=2e-5, abs=1e-6)
    after = dump_net(algo.nets["value_net"])
    for t in O.PARAM_ORDER:
        w = o_nets["value_net"][t]
        np.testing.assert_allclose(after[t], w, rtol=1e-4, atol=1e-5 * np.abs(w).max(), err_msg=t)


def test_reinforce_agent_loop():
    """recnn.nn.Reinforce wired like the Top-K notebook (cells 3-6): None on ordinary steps, a losses dict on policy
    steps, policy / targets move only then, everything finite."""
    torch.manual_seed(5)
    rng = np.random.default_rng(8)
    S, H, I, N = 52, 64, 300, 10
    policy = recnn_b200.nn.DiscreteActor(S, I, H)
    value = recnn_b200.nn.Critic(S, I, H, 54e-2)
    agent = recnn_b200.nn.Reinforce(policy, value).to(torch.device(DEV))
    policy = agent.nets["policy_net"]
    bw = torch.from_numpy(rng.normal(0, 0.3, (I, S)).astype(np.float32)).to(DEV)

    def select_action_corr(state, action, K, writer, step, **kwargs):
        beta = lambda s, action=None: torch.softmax(s @ bw.T, dim=1)       # noqa: E731
        return agent.nets["policy_net"]._select_action_with_TopK_correction(state, beta, action, K=K, writer=writer, step=step)

    policy.select_action = select_action_corr
    agent.params["reinforce"] = recnn_b200.nn.ChooseREINFORCE(recnn_b200.nn.ChooseREINFORCE.reinforce_with_TopK_correction)
    agent.params["K"] = 10
    agent.optimizers["policy_optimizer"] = recnn_b200.optim.Adam(policy.parameters(), lr=1e-3)
    agent.optimizers["value_optimizer"] = recnn_b200.optim.Adam(agent.nets["value_net"].parameters(), lr=1e-3)
    w0 = policy.linear2.weight.detach().clone()
    t0 = agent.nets["target_policy_net"].linear2.weight.detach().clone()
    v0 = agent.nets["value_net"].linear1.weight.detach().clone()
    out = []
    for i in range(21):
        b = {k: torch.from_numpy(v) for k, v in _dense_batch(rng, N, S, I).items()}
        out.append(agent.update(b))
        agent.step()
        if i == 9:
            assert torch.equal(w0, policy.linear2.weight.detach())                  # no policy update before step 10
            assert not torch.equal(v0, agent.nets["value_net"].linear1.weight.detach())   # the critic learns every step
            assert len(policy.rewards) == 10 and len(policy._saved) == 10
    assert [o is not None for o in out] == [i in (10, 20) for i in range(21)]
    for o in (out[10], out[20]):
        assert set(o) == {"value", "policy", "step"} and np.isfinite(o["value"]) and np.isfinite(o["policy"])
    assert not torch.equal(w0, policy.linear2.weight.detach())
    assert not torch.equal(t0, agent.nets["target_policy_net"].linear2.weight.detach())
    assert len(policy.rewards) == 0 and len(policy._saved) == 0


def test_errors():
    m = recnn_b200.nn.DiscreteActor(8, 20, 8).to(DEV)
    ch = recnn_b200.nn.ChooseREINFORCE()
    with pytest.raises(RuntimeError):
        ch.method(m, torch.zeros(0))                                   # nothing saved
    m._saved.append({"state": torch.zeros(2, 8, device=DEV), "action": torch.tensor([1, 20], device=DEV), "beta_log_prob": None})
    m.rewards += [torch.tensor(1.0), torch.tensor(2.0)]
    m._saved.append({"state": torch.zeros(2, 8, device=DEV), "action": torch.tensor([1, 2], device=DEV), "beta_log_prob": None})
    with pytest.raises(IndexError):
        ch(m, torch.optim.SGD(m.parameters(), lr=0.1))                 # action id == num_items
    with pytest.raises(TypeError):
        recnn_b200.nn.ChooseREINFORCE(lambda policy, returns: 0)(m, None)
    cpu = recnn_b200.nn.DiscreteActor(8, 20, 8)
    with pytest.raises(recnn_b200._lib.RecnnError):
        cpu(torch.zeros(2, 8))
