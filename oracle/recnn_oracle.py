"""CPU oracle for the RecNN DDPG/TD3 update hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy (float32) restatement of the reference algorithm with a
hand-derived backward pass.  It is the *checker* for the CUDA path; it is never
the thing shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from /root/reference (two import stubs for the off-path modules
matplotlib / torch_optimizer), runs ``batch_tensor_embeddings``, ``ddpg_update``
and ``td3_update`` on seeded inputs with replayed dropout masks, and stores the
results under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every
function below against those vectors.

Each function cites the reference lines it restates (paths relative to
/root/reference).  All arithmetic is float32 unless noted; scalars that the
reference keeps as Python floats (double) are kept double here too and cast at
the same point torch casts them.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------
# Data path
# ----------------------------------------------------------------------------
def rolling_windows(seq: np.ndarray, window: int) -> np.ndarray:
    """All length-``window`` sliding windows (recnn/data/utils.py:7-10)."""
    n = seq.shape[0] - window + 1
    idx = np.arange(n)[:, None] + np.arange(window)[None, :]
    return seq[idx]


def collate_users(users: list, frame_size: int):
    """recnn/data/utils.py:161-181: per-user windows of length frame_size+1,
    concatenated; ratings cast float64 -> float32 (``.float()``, :178)."""
    items = np.concatenate([rolling_windows(u["items"], frame_size + 1) for u in users], 0)
    rates = np.concatenate([rolling_windows(u["rates"], frame_size + 1) for u in users], 0)
    sizes = np.asarray([u["sizes"] for u in users], dtype=np.int64)
    uid = np.asarray([u["users"] for u in users], dtype=np.int64)
    return {"items": items.astype(np.int64), "ratings": rates.astype(F32),
            "sizes": sizes, "users": uid}


def collate_rows(users: list, frame_size: int, window_ids: np.ndarray):
    """Rows ``window_ids`` of the collate of ALL ``users`` (storage order): the fixed-size minibatch form
    of the device-resident feed.  Nothing new is computed: it is recnn/data/utils.py:161-181 over every
    user followed by row selection, and ``done`` is recnn/data/utils.py:70-71 of that full collate."""
    full = collate_users(users, frame_size)
    done = done_from_sizes(full["sizes"], frame_size, full["items"].shape[0])
    owner = np.repeat(full["users"], full["sizes"] - frame_size)       # user id of every row
    w = np.asarray(window_ids, dtype=np.int64)
    return {"items": full["items"][w], "ratings": full["ratings"][w], "done": done[w], "users": owner[w]}


def done_from_sizes(sizes: np.ndarray, frame_size: int, n_rows: int) -> np.ndarray:
    """recnn/data/utils.py:70-71: done[cumsum(sizes - F) - 1] = 1."""
    done = np.zeros(n_rows, dtype=F32)
    done[np.cumsum(sizes - frame_size) - 1] = 1.0
    return done


def frame_gather(table: np.ndarray, items: np.ndarray, ratings: np.ndarray,
                 sizes: np.ndarray, frame_size: int) -> dict:
    """Embedding gather + frame assembly (recnn/data/utils.py:51-81).

    table fp32[n_items, D]; items int64[N, F+1]; ratings fp32[N, F+1].
    Pure copy: results must be bit-exact.
    """
    n = ratings.shape[0]
    emb = table[items]                                   # :57
    state = np.concatenate([emb[:, :-1, :].reshape(n, -1), ratings[:, :-1]], 1)   # :60-65
    next_state = np.concatenate([emb[:, 1:, :].reshape(n, -1), ratings[:, 1:]], 1)  # :61-66
    return {
        "state": np.ascontiguousarray(state),
        "next_state": np.ascontiguousarray(next_state),
        "action": np.ascontiguousarray(emb[:, -1, :]),   # :67
        "reward": np.ascontiguousarray(ratings[:, -1]),  # :68
        "done": done_from_sizes(sizes, frame_size, n),   # :70-71
    }


# ----------------------------------------------------------------------------
# Networks.  A net is a dict {w1,b1,w2,b2,w3,b3}, nn.Linear layout ([out,in]).
# ----------------------------------------------------------------------------
PARAM_ORDER = ("w1", "b1", "w2", "b2", "w3", "b3")   # == nn.Module.parameters() order


# Smallest |pre-activation| seen by any kept hidden unit since the last reset.  ReLU makes the
# gradient a discontinuous function of the inputs: a unit whose pre-activation is within rounding
# error of 0 can have its gate decided differently by two correct fp32 implementations, which changes
# that unit's gradient by 1/N.  Test cases are chosen (oracle/find_seeds.py) so that this margin is
# far above fp32 rounding error, i.e. every gate decision is unambiguous.
GATE_MARGIN = {"min": float("inf")}


def reset_gate_margin():
    GATE_MARGIN["min"] = float("inf")


# Optional log of the AMBIGUOUS gates of a run: kept units (train-mode nets only: those are the ones that get a
# backward) whose |pre-activation| <= thresh.  Entries are (mask array, rows, cols): tests/_golden.py uses them to
# knock such units out of the replayed dropout masks, so that a full-size comparison has no gate decision that two
# correct fp32 implementations could take differently.
GATE_LOG = {"on": False, "thresh": 0.0, "hits": []}


def _hidden(x, w, b, mask):
    """relu(x W^T + b) then Dropout(p=.5) in train mode == * mask * 2
    (recnn/nn/models.py:66-69 / :208-211).  mask None <=> eval()."""
    z = x @ w.T + b
    az = np.abs(z) if mask is None else np.abs(z)[np.asarray(mask) != 0]
    if az.size:
        GATE_MARGIN["min"] = min(GATE_MARGIN["min"], float(az.min()))
    if GATE_LOG["on"] and mask is not None:
        rows, cols = np.nonzero((np.abs(z) <= F32(GATE_LOG["thresh"])) & (np.asarray(mask) != 0))
        if rows.size:
            GATE_LOG["hits"].append((mask, rows, cols))
    h = np.maximum(z, F32(0))
    if mask is not None:
        h = h * (mask.astype(F32) * F32(2.0))
    return h


def actor_forward(p: dict, state: np.ndarray, masks=None, tanh=False):
    """Actor.forward (recnn/nn/models.py:59-73)."""
    m1, m2 = masks if masks is not None else (None, None)
    h1 = _hidden(state, p["w1"], p["b1"], m1)
    h2 = _hidden(h1, p["w2"], p["b2"], m2)
    a = h2 @ p["w3"].T + p["b3"]
    if tanh:
        a = np.tanh(a)
    return a, (state, h1, h2, m1, m2)


def critic_forward(p: dict, state: np.ndarray, action: np.ndarray, masks=None):
    """Critic.forward (recnn/nn/models.py:205-213): cat([s, a], 1) first."""
    m1, m2 = masks if masks is not None else (None, None)
    x = np.concatenate([state, action], 1)
    h1 = _hidden(x, p["w1"], p["b1"], m1)
    h2 = _hidden(h1, p["w2"], p["b2"], m2)
    q = h2 @ p["w3"].T + p["b3"]
    return q, (x, h1, h2, m1, m2)


def _mlp_backward(p: dict, cache, d_out: np.ndarray, need_dx: bool):
    """Backward of the 3-layer MLP given dL/d(out).  Returns (grads, dx).
    The relu gate is h>0 <=> z>0 on kept units; dropped units have mask 0."""
    x, h1, h2, m1, m2 = cache
    g = {}
    g["w3"] = d_out.T @ h2
    g["b3"] = d_out.sum(0)
    dh2 = d_out @ p["w3"]
    gate2 = (h2 > 0).astype(F32) * (F32(2.0) if m2 is not None else F32(1.0))
    dz2 = dh2 * gate2
    g["w2"] = dz2.T @ h1
    g["b2"] = dz2.sum(0)
    dh1 = dz2 @ p["w2"]
    gate1 = (h1 > 0).astype(F32) * (F32(2.0) if m1 is not None else F32(1.0))
    dz1 = dh1 * gate1
    g["w1"] = dz1.T @ x
    g["b1"] = dz1.sum(0)
    dx = dz1 @ p["w1"] if need_dx else None
    return g, dx


# ----------------------------------------------------------------------------
# Losses / targets
# ----------------------------------------------------------------------------
def temporal_difference(reward, done, gamma, target):
    """recnn/nn/update/misc.py:6-7 (all [N,1])."""
    return reward + (F32(1.0) - done) * F32(gamma) * target


# ----------------------------------------------------------------------------
# Optimizers (torch.optim semantics, torch 2.11 single-tensor CPU path)
# ----------------------------------------------------------------------------
def make_optimizer(kind: str, **kw):
    o = {"kind": kind, "t": 0, "state": {}}
    if kind == "sgd":
        o.update(lr=kw.get("lr", 1e-3), momentum=kw.get("momentum", 0.0),
                 weight_decay=kw.get("weight_decay", 0.0))
    elif kind == "adam":
        o.update(lr=kw.get("lr", 1e-3), betas=kw.get("betas", (0.9, 0.999)),
                 eps=kw.get("eps", 1e-8), weight_decay=kw.get("weight_decay", 0.0))
    elif kind == "ranger":
        # torch_optimizer.Ranger defaults (the reference's default optimizer, recnn/nn/algo.py:84-89).  The package is
        # not vendored / installed: PARITY UNPINNED, restated from the published algorithm (RAdam + Lookahead).
        o.update(lr=kw.get("lr", 1e-3), alpha=kw.get("alpha", 0.5), k=kw.get("k", 6),
                 n_sma_threshold=kw.get("N_sma_threshhold", 5), betas=kw.get("betas", (0.95, 0.999)),
                 eps=kw.get("eps", 1e-5), weight_decay=kw.get("weight_decay", 0.0))
    else:
        raise ValueError(kind)
    return o


def _ranger_step(o: dict, p: dict, g: dict, t: int):
    """One torch_optimizer.Ranger step (RAdam with the N_sma threshold, then Lookahead every k steps)."""
    import math
    b1, b2 = o["betas"]
    beta2_t = b2 ** t
    n_sma_max = 2.0 / (1.0 - b2) - 1.0
    n_sma = n_sma_max - 2.0 * t * beta2_t / (1.0 - beta2_t)
    adaptive = n_sma > o["n_sma_threshold"]
    if adaptive:
        step_size = math.sqrt((1.0 - beta2_t) * (n_sma - 4.0) / (n_sma_max - 4.0) * (n_sma - 2.0) / n_sma
                              * n_sma_max / (n_sma_max - 2.0)) / (1.0 - b1 ** t)
    else:
        step_size = 1.0 / (1.0 - b1 ** t)
    for k in PARAM_ORDER:
        grad = g[k].astype(F32)
        st = o["state"].get(k)
        if st is None:
            st = {"m": np.zeros_like(p[k]), "v": np.zeros_like(p[k]), "slow": p[k].copy()}
            o["state"][k] = st
        st["v"] = (st["v"] * F32(b2) + (F32(1.0 - b2) * grad) * grad).astype(F32)
        st["m"] = (st["m"] * F32(b1) + F32(1.0 - b1) * grad).astype(F32)
        x = p[k]
        if o["weight_decay"] != 0.0:
            x = (x + F32(-o["weight_decay"] * o["lr"]) * x).astype(F32)
        if adaptive:
            denom = (np.sqrt(st["v"]) + F32(o["eps"])).astype(F32)
            x = (x + (F32(-step_size * o["lr"]) * st["m"]) / denom).astype(F32)
        else:
            x = (x + F32(-step_size * o["lr"]) * st["m"]).astype(F32)
        if t % o["k"] == 0:
            st["slow"] = (st["slow"] + F32(o["alpha"]) * (x - st["slow"])).astype(F32)
            x = st["slow"].copy()
        p[k] = x


def optimizer_step(o: dict, p: dict, g: dict):
    """torch.optim.SGD / torch.optim.Adam .step() (or torch_optimizer.Ranger) on every tensor of a net."""
    o["t"] += 1
    t = o["t"]
    if o["kind"] == "ranger":
        _ranger_step(o, p, g, t)
        return
    for k in PARAM_ORDER:
        grad = g[k].astype(F32)
        if o["weight_decay"] != 0.0:
            grad = grad + F32(o["weight_decay"]) * p[k]
        if o["kind"] == "sgd":
            if o["momentum"] != 0.0:
                buf = o["state"].get(k)
                buf = grad.copy() if buf is None else F32(o["momentum"]) * buf + grad
                o["state"][k] = buf
                grad = buf
            p[k] = (p[k] - F32(o["lr"]) * grad).astype(F32)
        else:
            b1, b2 = o["betas"]
            m, v = o["state"].get(k, (np.zeros_like(p[k]), np.zeros_like(p[k])))
            m = (m + F32(1.0 - b1) * (grad - m)).astype(F32)            # lerp_
            v = (v * F32(b2) + F32(1.0 - b2) * grad * grad).astype(F32)  # mul_, addcmul_
            o["state"][k] = (m, v)
            bc1 = 1.0 - b1 ** t
            bc2 = 1.0 - b2 ** t
            step_size = o["lr"] / bc1
            denom = (np.sqrt(v) / F32(bc2 ** 0.5) + F32(o["eps"])).astype(F32)
            p[k] = (p[k] - F32(step_size) * (m / denom)).astype(F32)


def soft_update(net: dict, target: dict, soft_tau: float):
    """recnn/utils/misc.py:1-5: t <- t*(1-tau) + p*tau (python-float scalars)."""
    for k in PARAM_ORDER:
        target[k] = (target[k] * F32(1.0 - soft_tau) + net[k] * F32(soft_tau)).astype(F32)


def clip_grad_quirk(g: dict, max_norm: float = -1.0):
    """torch.nn.utils.clip_grad_norm_(params, max_norm=-1, norm_type=1)
    (recnn/nn/update/ddpg.py:92, td3.py:133): with max_norm=-1 the "clip"
    coefficient is -1/(||g||_1 + 1e-6), clamped only from above, so every
    gradient is L1-normalised and sign-flipped."""
    total = F32(0)
    for k in PARAM_ORDER:
        total = F32(total + np.abs(g[k]).sum(dtype=F32))
    coef = F32(max_norm) / F32(total + F32(1e-6))
    coef = min(coef, F32(1.0))
    for k in PARAM_ORDER:
        g[k] = (g[k] * coef).astype(F32)
    return float(total)


# ----------------------------------------------------------------------------
# Update steps
# ----------------------------------------------------------------------------
def _col(x):
    return np.asarray(x, dtype=F32).reshape(-1, 1)     # get_base_batch unsqueeze(1), utils.py:269,273


def value_update(batch, params, nets, opts, masks, learn=True):
    """recnn/nn/update/misc.py:10-55.  masks = (m1, m2) for value_net."""
    s, a, s2 = batch["state"], batch["action"], batch["next_state"]
    r, d = _col(batch["reward"]), _col(batch["done"])
    a2, _ = actor_forward(nets["target_policy_net"], s2)                 # :28 (eval)
    q2, _ = critic_forward(nets["target_value_net"], s2, a2)             # :29
    y = temporal_difference(r, d, params["gamma"], q2)                    # :30-32
    y = np.clip(y, F32(params["min_value"]), F32(params["max_value"]))   # :33-35
    q, cache = critic_forward(nets["value_net"], s, a, masks)            # :37
    diff = q - y
    loss = F32(np.mean(diff * diff, dtype=F32))                           # :39
    grads = None
    if learn:
        d_q = (F32(2.0) * diff / F32(diff.size)).astype(F32)
        grads, _ = _mlp_backward(nets["value_net"], cache, d_q, need_dx=False)
        optimizer_step(opts["value_optimizer"], nets["value_net"], grads)  # :42-44
    return loss, {"next_action": a2, "target_value": q2, "expected_value": y,
                  "value": q, "value_grads": grads}


def _policy_loss_and_grads(policy, critic, s, pmasks, cmasks, want_grads):
    """-Q(s, pi(s)).mean() and, if asked, d/d(actor params)
    (recnn/nn/update/ddpg.py:78-91, td3.py:116-132)."""
    gen, pcache = actor_forward(policy, s, pmasks)
    qpi, ccache = critic_forward(critic, s, gen, cmasks)
    loss = F32(np.mean(-qpi, dtype=F32))
    grads = None
    if want_grads:
        d_q = np.full_like(qpi, F32(-1.0) / F32(qpi.size))
        _, dx = _mlp_backward(critic, ccache, d_q, need_dx=True)
        d_gen = np.ascontiguousarray(dx[:, s.shape[1]:])          # action slice of cat([s,a])
        grads, _ = _mlp_backward(policy, pcache, d_gen, need_dx=False)
    return loss, gen, grads


def ddpg_update(batch, params, nets, opts, masks, step, learn=True):
    """recnn/nn/update/ddpg.py:8-104.

    masks: six uint8/bool arrays [N,H] in the reference's drop_layer call order:
    (value m1,m2) -> (policy m1,m2) -> (value m1,m2 for the policy loss).
    nets/opts are mutated in place like the reference mutates its modules.
    """
    v_loss, dbg = value_update(batch, params, nets, opts, masks[0:2], learn)      # :63-73
    do_policy = bool(learn and step % params["policy_step"] == 0)                  # :89
    p_loss, gen, g = _policy_loss_and_grads(nets["policy_net"], nets["value_net"],
                                            batch["state"], masks[2:4], masks[4:6], do_policy)
    dbg["gen_action"] = gen
    if do_policy:
        dbg["policy_grad_l1"] = clip_grad_quirk(g)                                  # :92
        dbg["policy_grads"] = g
        optimizer_step(opts["policy_optimizer"], nets["policy_net"], g)             # :93
        soft_update(nets["value_net"], nets["target_value_net"], params["soft_tau"])    # :95-97
        soft_update(nets["policy_net"], nets["target_policy_net"], params["soft_tau"])  # :98-100
    return {"value": float(v_loss), "policy": float(p_loss), "step": step}, dbg


def td3_update(batch, params, nets, opts, masks, noise, step, learn=True):
    """recnn/nn/update/td3.py:8-150.

    masks: eight [N,H] arrays in call order: value_net1 (2), value_net2 (2),
    policy_net (2), value_net1 again for the policy loss (2).  ``noise`` is the
    raw N(0, noise_std) draw [N,A] (td3.py:74), clamped here (:77).
    """
    s, a, s2 = batch["state"], batch["action"], batch["next_state"]
    r, d = _col(batch["reward"]), _col(batch["done"])
    a2, _ = actor_forward(nets["target_policy_net"], s2)                           # :73
    nz = np.clip(noise.astype(F32), F32(-params["noise_clip"]), F32(params["noise_clip"]))  # :77
    a2 = a2 + nz                                                                    # :78
    q1t, _ = critic_forward(nets["target_value_net1"], s2, a2)                      # :81
    q2t, _ = critic_forward(nets["target_value_net2"], s2, a2)                      # :82
    y = temporal_difference(r, d, params["gamma"], np.minimum(q1t, q2t))            # :83-86 (no clamp)
    losses, dbg = {}, {"next_action": a2, "expected_value": y}
    for i, (name, oname) in enumerate((("value_net1", "value_optimizer1"),
                                       ("value_net2", "value_optimizer2"))):
        q, cache = critic_forward(nets[name], s, a, masks[2 * i:2 * i + 2])        # :88-89
        diff = q - y
        losses["value%d" % (i + 1)] = float(F32(np.mean(diff * diff, dtype=F32)))   # :91-93 MSELoss
        if learn:
            d_q = (F32(2.0) * diff / F32(diff.size)).astype(F32)
            g, _ = _mlp_backward(nets[name], cache, d_q, need_dx=False)
            optimizer_step(opts[oname], nets[name], g)                               # :96-102
    do_policy = bool(step % params["policy_update"] == 0 and learn)                  # :130
    p_loss, gen, g = _policy_loss_and_grads(nets["policy_net"], nets["value_net1"],
                                            s, masks[4:6], masks[6:8], do_policy)   # :116-127
    dbg["gen_action"] = gen
    if do_policy:
        dbg["policy_grad_l1"] = clip_grad_quirk(g)                                   # :133
        optimizer_step(opts["policy_optimizer"], nets["policy_net"], g)              # :134
        soft_update(nets["value_net1"], nets["target_value_net1"], params["soft_tau"])  # :136-138
        soft_update(nets["value_net2"], nets["target_value_net2"], params["soft_tau"])  # :139-141
        # NB: the reference never soft-updates target_policy_net in TD3.
    losses["policy"] = float(p_loss)
    losses["step"] = step
    return losses, dbg


# ----------------------------------------------------------------------------
# Deterministic synthetic inputs shared by tests / bench
# ----------------------------------------------------------------------------
def linear_init(rng: np.random.Generator, out_f: int, in_f: int, bound=None):
    """nn.Linear-shaped init: U(-1/sqrt(in), 1/sqrt(in)) unless ``bound``
    (the reference overrides linear3 with U(-init_w, init_w), models.py:56-57).
    Not bit-identical to torch's generator -- parity tests always copy the same
    arrays to both sides, so only the distribution matters."""
    b = bound if bound is not None else 1.0 / np.sqrt(in_f)
    w = rng.uniform(-b, b, size=(out_f, in_f)).astype(F32)
    bias = rng.uniform(-b, b, size=(out_f,)).astype(F32)
    return w, bias


def make_actor(rng, input_dim, action_dim, hidden, init_w=2e-1):
    w1, b1 = linear_init(rng, hidden, input_dim)
    w2, b2 = linear_init(rng, hidden, hidden)
    w3, b3 = linear_init(rng, action_dim, hidden, init_w)
    return {"w1": w1, "b1": b1, "w2": w2, "b2": b2, "w3": w3, "b3": b3}


def make_critic(rng, input_dim, action_dim, hidden, init_w=3e-5):
    w1, b1 = linear_init(rng, hidden, input_dim + action_dim)
    w2, b2 = linear_init(rng, hidden, hidden)
    w3, b3 = linear_init(rng, 1, hidden, init_w)
    return {"w1": w1, "b1": b1, "w2": w2, "b2": b2, "w3": w3, "b3": b3}


def copy_net(p: dict) -> dict:
    return {k: v.copy() for k, v in p.items()}


def synth_frames(rng, n_rows, n_items=26744, dim=128, frame_size=10, table=None):
    """SURVEY.md 8d S1 'iid' rows: items ~ U{0..n_items-1}, ratings ~ U{-4..5},
    one pseudo-user (sizes=[N+F]) so done[N-1]=1."""
    if table is None:
        table = rng.standard_normal((n_items, dim), dtype=F32)
    items = rng.integers(0, n_items, size=(n_rows, frame_size + 1), dtype=np.int64)
    ratings = rng.integers(-4, 6, size=(n_rows, frame_size + 1)).astype(F32)
    sizes = np.asarray([n_rows + frame_size], dtype=np.int64)
    return table, items, ratings, sizes


def synth_masks(rng, count, n_rows, hidden):
    return [rng.integers(0, 2, size=(n_rows, hidden), dtype=np.uint8) for _ in range(count)]


# ----------------------------------------------------------------------------
# Serving: nearest-item retrieval (examples/streamlit_demo.py:189-215 faiss IndexFlatL2 / IndexFlatIP / cosine;
# recnn/data/db_con.py:45-56 MilvusConnection.search).  Brute force in float64, stable order (ties -> smaller id).
# ----------------------------------------------------------------------------
def retrieve_topk(queries: np.ndarray, table: np.ndarray, k: int, metric: str = "L2"):
    """-> (ids int64 [n,k], dist float64 [n,k]) best first.  L2: squared distance (faiss IndexFlatL2 / Milvus L2),
    IP: inner product, COS: cosine similarity (IndexFlatIP over L2-normalised rows, streamlit_demo.py:196-202)."""
    q = np.asarray(queries, dtype=np.float64)
    t = np.asarray(table, dtype=np.float64)
    if metric == "L2":
        d = (q * q).sum(1)[:, None] + (t * t).sum(1)[None, :] - 2.0 * (q @ t.T)
        d = np.maximum(d, 0.0)
        key = d
    elif metric == "IP":
        d = q @ t.T
        key = -d
    elif metric == "COS":
        d = (q @ t.T) / np.maximum(np.linalg.norm(q, axis=1), 1e-30)[:, None] / np.maximum(np.linalg.norm(t, axis=1), 1e-30)[None, :]
        key = -d
    else:
        raise ValueError(metric)
    ids = np.argsort(key, axis=1, kind="stable")[:, :k].astype(np.int64)
    return ids, np.take_along_axis(d, ids, axis=1)
