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
