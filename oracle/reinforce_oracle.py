"""CPU oracle for the REINFORCE (Top-K off-policy correction) policy side  --  TEST INFRASTRUCTURE ONLY.

numpy restatement of recnn/nn/models.py:76-184 (DiscreteActor) and recnn/nn/update/reinforce.py:10-65
(ChooseREINFORCE) with a hand-derived backward pass.  Checker for ``recnn_b200/csrc/reinforce.cuh``; never shipped or
measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may import it.

Parity status: PINNED.  ``oracle/make_reinforce_golden.py`` imports the unmodified reference, runs
``DiscreteActor._select_action`` / ``_select_action_with_correction`` / ``_select_action_with_TopK_correction`` for several
env steps (the Categorical draws are recorded by wrapping ``recnn.nn.models.Categorical``; no reference source is
touched) and ``ChooseREINFORCE(method)(policy, SGD, learn=True)``, and stores inputs, draws, log-probs, corrections,
lambda_K, returns, loss, gradients and stepped parameters in ``tests/golden/reinforce_*.npz``;
``tests/test_reinforce_oracle.py`` checks every function below against them.

Arithmetic is float64 where noted (the oracle is the checker: it should be closer to the exact answer than either
fp32 implementation), float32 at the interfaces.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
EPS = float(np.finfo(np.float32).eps)       # torch.distributions clamps probabilities to [eps, 1 - eps]

BASIC, CORRECTED, TOPK = 0, 1, 2
METHODS = {"basic_reinforce": BASIC, "reinforce_with_correction": CORRECTED, "reinforce_with_TopK_correction": TOPK}


def make_discrete_actor(rng: np.random.Generator, input_dim: int, num_items: int, hidden: int) -> dict:
    """nn.Linear default init (models.py:80-81; init_w is unused by the reference)."""
    def lin(out_f, in_f):
        b = 1.0 / np.sqrt(in_f)
        return (rng.uniform(-b, b, (out_f, in_f)).astype(F32), rng.uniform(-b, b, (out_f,)).astype(F32))
    w1, b1 = lin(hidden, input_dim)
    w2, b2 = lin(num_items, hidden)
    return {"w1": w1, "b1": b1, "w2": w2, "b2": b2}


def discrete_forward(p: dict, state: np.ndarray, dtype=np.float64):
    """models.py:95-99: softmax(linear2(relu(linear1(x)))).  Returns (probs, hidden)."""
    x = state.astype(dtype)
    h = np.maximum(x @ p["w1"].astype(dtype).T + p["b1"].astype(dtype), 0)
    z = h @ p["w2"].astype(dtype).T + p["b2"].astype(dtype)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True), h


def categorical_log_prob(probs: np.ndarray, action: np.ndarray) -> np.ndarray:
    """Categorical(probs).log_prob(action): probs normalised by their row sum, clamped to [eps, 1-eps]
    (torch/distributions/categorical.py __init__, utils.probs_to_logits)."""
    pr = probs.astype(np.float64)
    pa = pr[np.arange(pr.shape[0]), action] / pr.sum(axis=1)
    return np.log(np.clip(pa, EPS, 1.0 - EPS))


def categorical_sample(probs: np.ndarray, uniforms: np.ndarray):
    """Inverse-CDF draw: first j with cumsum(probs)[j] > u * sum(probs) (and probs[j] > 0).  Returns
    (action, log_prob, margin): margin = distance of u*sum from the nearest interval end, relative to the sum --
    tests skip rows whose margin is below the fp32 summation noise."""
    pr = probs.astype(np.float64)
    c = np.cumsum(pr, axis=1)
    tot = c[:, -1]
    tgt = uniforms.astype(np.float64) * tot
    n = pr.shape[0]
    act = np.empty(n, np.int64)
    margin = np.empty(n, np.float64)
    for r in range(n):
        j = int(np.searchsorted(c[r], tgt[r], side="right"))
        while j < pr.shape[1] - 1 and pr[r, j] <= 0:
            j += 1
        j = min(j, pr.shape[1] - 1)
        act[r] = j
        lo = c[r, j - 1] if j > 0 else 0.0
        margin[r] = min(tgt[r] - lo, c[r, j] - tgt[r]) / tot[r]
    return act, categorical_log_prob(probs, act), margin


def normalised_returns(rewards) -> np.ndarray:
    """reinforce.py:44-52: discounted returns over the saved env steps (the discount is the literal 0.99 of
    reinforce.py:48, not params["gamma"]), normalised with the UNBIASED std (torch.Tensor.std default) + 1e-4;
    float32 like torch.tensor(list of float32 scalars)."""
    R = F32(0.0)
    out = []
    for r in list(rewards)[::-1]:
        R = F32(F32(r) + F32(F32(0.99) * R))
        out.insert(0, R)
    ret = np.asarray(out, F32)
    mean = ret.mean(dtype=np.float64)
    std = np.sqrt(((ret.astype(np.float64) - mean) ** 2).sum() / (len(ret) - 1)) if len(ret) > 1 else np.nan
    return ((ret - F32(mean)) / F32(F32(std) + F32(0.0001))).astype(F32)


def row_terms(pa: np.ndarray, beta_logp, ret: np.ndarray, method: int, K: int):
    """Per-row loss L and dL/d(log pi[a]) for the three ChooseREINFORCE methods (reinforce.py:16-44 with
    models.py:150 and :168-171: corr and lambda_K carry gradient)."""
    p = np.clip(pa.astype(np.float64), EPS, 1.0 - EPS)
    inside = (pa > EPS) & (pa < 1.0 - EPS)
    lp = np.log(p)
    R = ret.astype(np.float64)
    if method == BASIC:
        L, g = -lp * R, -R
    else:
        c = p / np.exp(np.asarray(beta_logp, np.float64))
        if method == CORRECTED:
            L, g = c * -lp * R, -R * c * (lp + 1.0)
        else:
            q = 1.0 - p
            lam = K * q ** (K - 1)
            dlam = -K * (K - 1) * q ** (K - 2) * p if K > 1 else np.zeros_like(p)
            L, g = lam * c * -lp * R, -R * c * (dlam * lp + lam * (lp + 1.0))
    return L, np.where(inside, g, 0.0), lp


def reinforce_policy_grad(p: dict, state: np.ndarray, action: np.ndarray, beta_logp, ret: np.ndarray,
                          method: int, K: int = 10):
    """Loss (sum over the saved rows) and gradient of every DiscreteActor parameter.  float64 inside."""
    probs, h = discrete_forward(p, state)
    n = state.shape[0]
    pa = probs[np.arange(n), action]
    L, g, lp = row_terms(pa, beta_logp, ret, method, K)
    dz = -probs * g[:, None]
    dz[np.arange(n), action] += g
    gw2 = dz.T @ h
    gb2 = dz.sum(0)
    dh = (dz @ p["w2"].astype(np.float64)) * (h > 0)
    gw1 = dh.T @ state.astype(np.float64)
    gb1 = dh.sum(0)
    return float(L.sum()), {"w1": gw1, "b1": gb1, "w2": gw2, "b2": gb2}, {"probs": probs, "log_prob": lp, "row_loss": L}


def correction(pi_logp, beta_logp):
    """models.py:150 / :165."""
    return np.exp(np.asarray(pi_logp, np.float64)) / np.exp(np.asarray(beta_logp, np.float64))


def lambda_k(pi_logp, K: int):
    """models.py:168."""
    return K * (1.0 - np.exp(np.asarray(pi_logp, np.float64))) ** (K - 1)


def value_update(batch, params, nets, opts, masks, learn=True):
    """The critic half of reinforce_update (reinforce.py:92-102 -> misc.py:10-55) with a DiscreteActor as the target
    policy: next_action = target_policy_net(next_state) are PROBABILITIES [N, num_items], the batch action is the dense
    one-hot.  Built from the DDPG oracle's primitives (oracle/recnn_oracle.py)."""
    from oracle import recnn_oracle as O
    s, a, s2 = batch["state"], batch["action"], batch["next_state"]
    r, d = O._col(batch["reward"]), O._col(batch["done"])
    a2 = discrete_forward(nets["target_policy_net"], s2, dtype=np.float32)[0].astype(F32)     # misc.py:28
    q2, _ = O.critic_forward(nets["target_value_net"], s2, a2)                                 # :29
    y = O.temporal_difference(r, d, params["gamma"], q2)                                       # :30-32
    y = np.clip(y, F32(params["min_value"]), F32(params["max_value"]))                        # :33-35
    q, cache = O.critic_forward(nets["value_net"], s, a, masks)                                # :37
    diff = q - y
    loss = F32(np.mean(diff * diff, dtype=F32))                                                # :39
    if learn:
        d_q = (F32(2.0) * diff / F32(diff.size)).astype(F32)
        grads, _ = O._mlp_backward(nets["value_net"], cache, d_q, need_dx=False)
        O.optimizer_step(opts["value_optimizer"], nets["value_net"], grads)                    # :42-44
    return loss, {"next_action": a2, "expected_value": y, "value": q}


# ----------------------------------------------------------------------------------------------------------------------
# Vocabulary-parallel formulation (BASELINE configs[4]: 1M items over 8 GPUs) -- DESIGN STUDY, no CUDA counterpart yet.
# The item dimension of linear2 (and of the softmax) is sharded over W ranks; linear1 is replicated.  The functions below
# restate the policy gradient as W per-rank computations plus the three exchanges a device implementation needs, and
# tests/test_reinforce_oracle.py checks them against the unsharded oracle:
#   exchange 1 (all-gather, 3 floats per row and rank): local max m_r, local sum s_r = sum exp(z - m_r), and the logit of
#               the drawn action (owner rank only, 0 elsewhere)
#   exchange 2 (all-reduce, [rows, hidden]): dh = sum over ranks of dz_r @ W2_r
# Everything else (dW2_r, db2_r of the shard; dW1, db1 replicated from the all-reduced dh) is rank-local.
def shard_policy(p: dict, world: int):
    """Split a DiscreteActor's linear2 rows into `world` contiguous shards (the last may be shorter)."""
    items = p["w2"].shape[0]
    per = -(-items // world)
    shards = []
    for r in range(world):
        lo, hi = min(r * per, items), min((r + 1) * per, items)
        shards.append({"w1": p["w1"], "b1": p["b1"], "w2": p["w2"][lo:hi], "b2": p["b2"][lo:hi], "offset": lo})
    return shards


def sharded_policy_grad(shards: list, state: np.ndarray, action: np.ndarray, beta_logp, ret: np.ndarray,
                        method: int, K: int = 10):
    """Loss and per-rank gradients of the vocabulary-sharded policy; see the exchanges above."""
    n = state.shape[0]
    x = state.astype(np.float64)
    local = []
    for sh in shards:                                                        # ---- rank-local forward
        h = np.maximum(x @ sh["w1"].astype(np.float64).T + sh["b1"].astype(np.float64), 0)
        z = h @ sh["w2"].astype(np.float64).T + sh["b2"].astype(np.float64)
        cnt = z.shape[1]
        m = z.max(axis=1) if cnt else np.full(n, -np.inf)
        s = np.exp(z - m[:, None]).sum(axis=1) if cnt else np.zeros(n)
        mine = (action >= sh["offset"]) & (action < sh["offset"] + cnt)
        za = np.where(mine, z[np.arange(n), np.clip(action - sh["offset"], 0, max(cnt - 1, 0))] if cnt else 0.0, 0.0)
        local.append({"h": h, "z": z, "m": m, "s": s, "za": za, "mine": mine})
    # ---- exchange 1: every rank now holds (m_r, s_r, za_r) of all ranks
    M = np.max([l["m"] for l in local], axis=0)
    S = np.sum([l["s"] * np.exp(l["m"] - M) for l in local], axis=0)
    za = np.sum([l["za"] for l in local], axis=0)
    pa = np.exp(za - M) / S                                                  # pi[a], identical on every rank
    L, g, _ = row_terms(pa, beta_logp, ret, method, K)
    grads, dh = [], np.zeros_like(local[0]["h"])
    for sh, l in zip(shards, local):                                         # ---- rank-local backward
        probs = np.exp(l["z"] - M[:, None]) / S[:, None]
        dz = -probs * g[:, None]
        rows = np.nonzero(l["mine"])[0]
        dz[rows, action[rows] - sh["offset"]] += g[rows]
        grads.append({"w2": dz.T @ l["h"], "b2": dz.sum(0)})
        dh += dz @ sh["w2"].astype(np.float64)                               # ---- exchange 2: all-reduce of dh
    dh = dh * (local[0]["h"] > 0)
    for gr in grads:                                                         # replicated layer 1
        gr["w1"] = dh.T @ x
        gr["b1"] = dh.sum(0)
    return float(L.sum()), grads
