"""The numpy oracle (oracle/recnn_oracle.py) against the golden vectors that
oracle/make_golden.py produced by running the unmodified reference."""
import numpy as np
import pytest

from oracle import recnn_oracle as O
from tests._golden import load_golden, run_oracle_case, compare_with_golden


def test_gather_bit_exact_vs_reference():
    g = load_golden("gather.npz")
    users = []
    for i in range(3):
        users.append({"items": g["user%d.items" % i], "rates": g["user%d.rates" % i],
                      "sizes": len(g["user%d.items" % i]), "users": int(g["user%d.id" % i])})
    frame = int(g["frame_size"])
    col = O.collate_users(users, frame)
    out = O.frame_gather(g["table"], col["items"], col["ratings"], col["sizes"], frame)
    for k in ("state", "next_state", "action", "reward", "done"):
        assert out[k].dtype == np.float32
        assert out[k].shape == g["out." + k].shape
        assert np.array_equal(out[k].view(np.uint32), g["out." + k].view(np.uint32)), k
    assert np.array_equal(col["sizes"], g["out.sizes"])
    assert np.array_equal(col["users"], g["out.users"])
    # within a user next_state[i] == state[i+1]  
    assert np.array_equal(out["next_state"][0], out["state"][1])


@pytest.mark.parametrize("case", ["tiny", "canon"])
@pytest.mark.parametrize("opt", ["adam", "sgd"])
def test_ddpg_oracle_vs_reference(case, opt):
    gold = load_golden("ddpg_%s_%s.npz" % (case, opt))
    got = run_oracle_case(case, "ddpg", opt)
    compare_with_golden(got, gold)


@pytest.mark.parametrize("case", ["tiny", "canon"])
@pytest.mark.parametrize("opt", ["adam", "sgd"])
def test_td3_oracle_vs_reference(case, opt):
    gold = load_golden("td3_%s_%s.npz" % (case, opt))
    got = run_oracle_case(case, "td3", opt, golden=gold)
    compare_with_golden(got, gold, check_grads=False)


def test_quirks():
    """Load-bearing quirks."""
    rng = np.random.default_rng(0)
    g = {k: rng.standard_normal(s).astype(np.float32) for k, s in
         zip(O.PARAM_ORDER, [(4, 3), (4,), (4, 4), (4,), (2, 4), (2,)])}
    before = {k: v.copy() for k, v in g.items()}
    total = O.clip_grad_quirk(g)
    l1 = sum(np.abs(v).sum() for v in g.values())
    assert abs(l1 - 1.0) < 1e-5                      # L1-normalised
    for k in g:                                       # sign flipped
        assert np.all(np.sign(g[k]) == -np.sign(before[k]))
    assert total > 0


def test_oracle_ranger_radam_half_matches_torch_radam():
    """torch_optimizer.Ranger is absent (parity unpinned); its RAdam half is the same update as
    torch.optim.RAdam(decoupled_weight_decay=True), which IS available: with Lookahead disabled (k beyond the run)
    the oracle's Ranger must track it over the rectification switch (steps 1..5 un-rectified, then rectified)."""
    import torch
    from oracle import recnn_oracle as O
    rng = np.random.default_rng(0)
    p0 = {k: rng.standard_normal(s).astype(np.float32) for k, s in
          zip(O.PARAM_ORDER, [(8, 5), (8,), (8, 8), (8,), (3, 8), (3,)])}
    p = {k: v.copy() for k, v in p0.items()}
    tp = [torch.nn.Parameter(torch.from_numpy(p0[k].copy())) for k in O.PARAM_ORDER]
    topt = torch.optim.RAdam(tp, lr=1e-2, betas=(0.95, 0.999), eps=1e-5, weight_decay=1e-2, decoupled_weight_decay=True)
    o = O.make_optimizer("ranger", lr=1e-2, weight_decay=1e-2, k=10 ** 9)
    for step in range(12):
        g = {k: rng.standard_normal(p0[k].shape).astype(np.float32) for k in O.PARAM_ORDER}
        for q, k in zip(tp, O.PARAM_ORDER):
            q.grad = torch.from_numpy(g[k].copy())
        topt.step()
        O.optimizer_step(o, p, g)
        for q, k in zip(tp, O.PARAM_ORDER):
            np.testing.assert_allclose(p[k], q.detach().numpy(), rtol=2e-5, atol=2e-6, err_msg="%s step %d" % (k, step))


def test_oracle_ranger_lookahead_half():
    """Every k-th step the weights are pulled half way (alpha) back to the slow copy, which starts at the initial weights."""
    from oracle import recnn_oracle as O
    rng = np.random.default_rng(1)
    p = {k: rng.standard_normal((4, 4) if k.startswith("w") else (4,)).astype(np.float32) for k in O.PARAM_ORDER}
    init = {k: v.copy() for k, v in p.items()}
    fast = {k: v.copy() for k, v in p.items()}
    o = O.make_optimizer("ranger", lr=1e-2, k=3, alpha=0.5)
    o_fast = O.make_optimizer("ranger", lr=1e-2, k=10 ** 9)
    for step in range(3):
        g = {k: rng.standard_normal(p[k].shape).astype(np.float32) for k in O.PARAM_ORDER}
        O.optimizer_step(o, p, g)
        O.optimizer_step(o_fast, fast, g)
    for k in O.PARAM_ORDER:
        np.testing.assert_allclose(p[k], init[k] + 0.5 * (fast[k] - init[k]), rtol=1e-6, atol=1e-7)
