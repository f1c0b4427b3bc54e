"""Device-resident FrameEnv feed (recnn_b200/data/feed.py, csrc/feed.cu).

CPU part: the numpy oracle of the collate against the golden fixture produced by the unmodified
reference (tests/golden/collate.npz, oracle/make_golden.py:run_collate_case), and the host-side
planning logic of ``HistoryCSR``.  GPU part: the window-gather kernels, bit-exact against the golden
fixture and the live oracle, through the C ABI; the full feed -> update path against a host-collated
batch."""
import numpy as np
import pytest
import torch

import recnn_b200
from recnn_b200 import _lib
from recnn_b200.data.feed import HistoryCSR, DeviceFrameFeed
from oracle import recnn_oracle as O
from tests._golden import load_golden


def golden_users(g):
    users = []
    for i in range(int(g["n_users"])):
        users.append({"items": g["user%d.items" % i], "rates": g["user%d.rates" % i],
                      "sizes": len(g["user%d.items" % i]), "users": int(g["user%d.id" % i])})
    return users


def csr_of(users, frame):
    return HistoryCSR([u["users"] for u in users], [u["items"] for u in users], [u["rates"] for u in users], frame)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else a.dtype)


# =============================================================================== CPU: oracle vs reference
@pytest.mark.parametrize("tag", ["all", "mini"])
def test_oracle_collate_bit_exact_vs_reference(tag):
    g = load_golden("collate.npz")
    users = golden_users(g)
    frame = int(g["frame_size"])
    sel = list(range(len(users))) if tag == "all" else g["minibatch"].tolist()
    col = O.collate_users([users[i] for i in sel], frame)
    assert col["items"].dtype == np.int64 and col["ratings"].dtype == np.float32
    assert np.array_equal(col["items"], g[tag + ".items"])
    assert np.array_equal(bits(col["ratings"]), bits(g[tag + ".ratings"]))      # incl. the f64 -> f32 rounding
    assert np.array_equal(col["sizes"], g[tag + ".sizes"])
    assert np.array_equal(col["users"], g[tag + ".users"])
    done = O.done_from_sizes(col["sizes"], frame, col["items"].shape[0])
    assert np.array_equal(done, g[tag + ".done"])
    # some ratings of the fixture are not fp32-representable, so the cast is really exercised
    raw = np.concatenate([u["rates"] for u in users])
    assert np.any(raw.astype(np.float32).astype(np.float64) != raw)


def test_oracle_rows_by_window_id_are_rows_of_the_reference_collate():
    g = load_golden("collate.npz")
    users = golden_users(g)
    frame = int(g["frame_size"])
    n = g["all.items"].shape[0]
    w = np.random.default_rng(5).integers(0, n, size=400)
    rows = O.collate_rows(users, frame, w)
    assert np.array_equal(rows["items"], g["all.items"][w])
    assert np.array_equal(bits(rows["ratings"]), bits(g["all.ratings"][w]))
    assert np.array_equal(rows["done"], g["all.done"][w])
    owner = np.repeat(g["all.users"], g["all.sizes"] - frame)
    assert np.array_equal(rows["users"], owner[w])


# =============================================================================== CPU: host planning logic
def test_history_csr_layout_and_plans():
    g = load_golden("collate.npz")
    users = golden_users(g)
    frame = int(g["frame_size"])
    csr = csr_of(users, frame)
    assert csr.n_users == len(users) and csr.n_windows == g["all.items"].shape[0]
    assert np.array_equal(csr.lengths, g["all.sizes"])
    assert csr.ratings.dtype == np.float32 and csr.items.dtype == np.int64
    for i, u in enumerate(users):
        lo, hi = csr.offsets[i], csr.offsets[i + 1]
        assert np.array_equal(csr.items[lo:hi], u["items"])
        assert np.array_equal(bits(csr.ratings[lo:hi]), bits(u["rates"].astype(np.float32)))
    sel = g["minibatch"]
    row_offsets, n_rows = csr.plan_users(sel)
    assert n_rows == g["mini.items"].shape[0]
    assert np.array_equal(np.diff(row_offsets), g["mini.sizes"] - frame)
    # done positions of the reference = last row of every user = row_offsets[1:] - 1
    assert np.array_equal(np.nonzero(g["mini.done"])[0], row_offsets[1:] - 1)
    with pytest.raises(IndexError):
        csr.plan_users([0, len(users)])
    with pytest.raises(ValueError):
        HistoryCSR([1], [np.arange(5)], [np.zeros(5)], frame).plan_users([0])     # 5 < frame_size interactions
    with pytest.raises(ValueError):
        HistoryCSR([1], [np.arange(12)], [np.zeros(11)], frame)


@pytest.mark.parametrize("batch_size,drop_last", [(4, False), (4, True), (14, False), (25, False), (1, False)])
def test_epoch_plan_matches_dataloader_batching(batch_size, drop_last):
    g = load_golden("collate.npz")
    users = golden_users(g)
    frame = int(g["frame_size"])
    csr = csr_of(users, frame)
    perm = torch.randperm(len(users), generator=torch.Generator().manual_seed(3)).numpy()
    starts, counts, flat, row_starts, n_rows = csr.plan_epoch(perm, batch_size, drop_last)
    # torch's own batching of the same permutation
    want = list(torch.utils.data.BatchSampler(perm.tolist(), batch_size, drop_last))
    assert len(want) == len(starts)
    for b, idx in enumerate(want):
        assert perm[starts[b]:starts[b] + counts[b]].tolist() == idx
        col = O.collate_users([users[i] for i in idx], frame)
        plan = flat[row_starts[b]:row_starts[b] + counts[b] + 1]
        assert plan[0] == 0 and n_rows[b] == col["items"].shape[0] == plan[-1]
        assert np.array_equal(np.diff(plan), col["sizes"] - frame)


def test_feed_refuses_cpu():
    g = load_golden("collate.npz")
    csr = csr_of(golden_users(g), int(g["frame_size"]))
    with pytest.raises(_lib.RecnnError):
        DeviceFrameFeed(csr, torch.zeros(500, 4), device="cpu")


# =============================================================================== GPU: kernels through the C ABI
def _feed(dim=4, n_items=500, device="cuda:0"):
    g = load_golden("collate.npz")
    users = golden_users(g)
    frame = int(g["frame_size"])
    table = np.random.default_rng(78).standard_normal((n_items, dim), dtype=np.float32)
    return g, users, frame, table, DeviceFrameFeed(csr_of(users, frame), torch.from_numpy(table), device)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["all", "mini"])
def test_window_gather_users_bit_exact_vs_reference(tag):
    g, users, frame, table, feed = _feed()
    sel = list(range(len(users))) if tag == "all" else g["minibatch"].tolist()
    before = _lib.lib().recnn_b200_launch_count()
    b = feed.batch(sel, check=True)
    assert _lib.lib().recnn_b200_launch_count() == before + 1          # one kernel per minibatch
    assert b["items"].dtype == torch.int64 and b["ratings"].dtype == torch.float32
    assert np.array_equal(b["items"].cpu().numpy(), g[tag + ".items"])
    assert np.array_equal(bits(b["ratings"].cpu().numpy()), bits(g[tag + ".ratings"]))
    assert np.array_equal(b["done"].cpu().numpy(), g[tag + ".done"])
    assert np.array_equal(b["sizes"].cpu().numpy(), g[tag + ".sizes"])
    assert np.array_equal(b["users"].numpy(), g[tag + ".users"])
    assert np.array_equal(b["meta"]["sizes"].numpy(), g[tag + ".sizes"])


@pytest.mark.gpu
def test_window_gather_ids_bit_exact_vs_reference_rows():
    g, users, frame, table, feed = _feed()
    n = g["all.items"].shape[0]
    w = np.concatenate([np.arange(n), np.random.default_rng(9).integers(0, n, size=5000)])   # every window + random
    b = feed.windows(torch.from_numpy(w), check=True)
    assert np.array_equal(b["items"].cpu().numpy(), g["all.items"][w])
    assert np.array_equal(bits(b["ratings"].cpu().numpy()), bits(g["all.ratings"][w]))
    assert np.array_equal(b["done"].cpu().numpy(), g["all.done"][w])
    owner_pos = np.repeat(np.arange(len(users)), g["all.sizes"] - frame)
    assert np.array_equal(b["user_positions"].cpu().numpy(), owner_pos[w])
    with pytest.raises(IndexError):
        feed.windows(torch.tensor([0, n]), check=True)            # id == n_windows is out of range
    with pytest.raises(IndexError):
        feed.windows(torch.tensor([-1]), check=True)


@pytest.mark.gpu
def test_window_gather_detects_a_plan_that_disagrees_with_the_histories():
    g, users, frame, table, feed = _feed()
    L = _lib.lib()
    dev = feed.device
    pos = torch.tensor([2, 0], dtype=torch.int64, device=dev)
    wrong = torch.tensor([0, 5, 6], dtype=torch.int64, device=dev)     # user 2 has 20 windows, not 5
    items = torch.full((6, frame + 1), 7, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.recnn_window_gather_users(feed.hist_items.data_ptr(), feed.hist_ratings.data_ptr(),
                                           feed.hist_offsets.data_ptr(), feed.csr.n_users, pos.data_ptr(),
                                           wrong.data_ptr(), 2, frame, 6, items.data_ptr(), None, None, None,
                                           err.data_ptr(), _lib.stream_ptr(dev)))
    assert int(err.item()) == 1
    assert int(items[:5].abs().sum().item()) == 0                      # affected rows are zero-filled
    assert np.array_equal(items[5].cpu().numpy(), users[0]["items"][:frame + 1])   # user 0 has exactly 1 window
    # argument validation happens on the host side of the ABI
    assert L.recnn_window_gather_users(None, None, None, 1, None, None, 1, frame, 1, None, None, None, None,
                                       None, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("batch_size", [4, 25])
def test_epoch_yields_the_reference_minibatches(batch_size):
    g, users, frame, table, feed = _feed()
    gen = torch.Generator().manual_seed(11)
    perm = torch.randperm(len(users), generator=torch.Generator().manual_seed(11)).tolist()
    seen = 0
    for b, idx in zip(feed.epoch(batch_size, generator=gen),
                      torch.utils.data.BatchSampler(perm, batch_size, False)):
        col = O.collate_users([users[i] for i in idx], frame)
        assert np.array_equal(b["items"].cpu().numpy(), col["items"])
        assert np.array_equal(bits(b["ratings"].cpu().numpy()), bits(col["ratings"]))
        assert np.array_equal(b["done"].cpu().numpy(), O.done_from_sizes(col["sizes"], frame, col["items"].shape[0]))
        assert np.array_equal(b["users"].numpy(), col["users"])
        seen += len(idx)
    assert seen == len(users)


@pytest.mark.gpu
def test_feed_embed_equals_reference_state_frames():
    """feed.batch -> feed.embed == prepare_batch_static_size + batch_tensor_embeddings (gather.npz users)."""
    g = load_golden("gather.npz")
    users = [{"items": g["user%d.items" % i], "rates": g["user%d.rates" % i], "users": int(g["user%d.id" % i])}
             for i in range(3)]
    frame = int(g["frame_size"])
    feed = DeviceFrameFeed(csr_of(users, frame), torch.from_numpy(g["table"]), "cuda:0")
    out = feed.embed(feed.batch([0, 1, 2], check=True))
    for k in ("state", "next_state", "action", "reward", "done"):
        assert np.array_equal(bits(out[k].cpu().numpy()), bits(g["out." + k])), k


@pytest.mark.gpu
def test_sample_is_constant_size_and_reproducible():
    g, users, frame, table, feed = _feed()
    gen = torch.Generator(device=feed.device).manual_seed(5)
    a = feed.sample(256, generator=gen)
    gen.manual_seed(5)
    b = feed.sample(256, generator=gen)
    assert a["items"].shape == (256, frame + 1) and torch.equal(a["window_ids"], b["window_ids"])
    assert torch.equal(a["items"], b["items"])
    rows = O.collate_rows(users, frame, a["window_ids"].cpu().numpy())
    assert np.array_equal(a["items"].cpu().numpy(), rows["items"])
    assert np.array_equal(a["done"].cpu().numpy(), rows["done"])


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["ddpg", "td3"])
def test_update_from_the_feed_equals_update_from_a_host_collated_batch(algo):
    """The whole seam: FrameEnv.device_feed().batch(...) -> ddpg_update must train exactly like the
    host-collated frame batch of the same users (bit-identical losses and weights: same kernels, same
    inputs), i.e. the feed changes where the minibatch is built, not what is computed."""
    from oracle import cases as C
    from tests._cuda import build_nets, build_optimizers, dump_net
    spec = C.CASES["tiny"]
    inp = C.make_inputs(spec, algo)
    frame, dev = spec["frame"], torch.device("cuda:0")
    n_items = inp["table"].shape[0]
    rng = np.random.default_rng(21)
    user_dict = {}
    for uid, length in ((5, frame + 1), (9, frame + 7), (2, frame + 20), (40, frame + 3)):
        user_dict[uid] = {"items": rng.integers(0, n_items, size=length, dtype=np.int64),
                          "ratings": rng.integers(-4, 6, size=length).astype(np.float64)}
    env = recnn_b200.data.FrameEnv.from_user_dict(torch.from_numpy(inp["table"]), user_dict, frame_size=frame,
                                                  batch_size=4, test_size=0.0)
    feed = env.device_feed(dev)
    order = [2, 0, 3, 1]
    ds = env.base.train_user_dataset
    host = recnn_b200.data.prepare_batch_static_size([ds[i] for i in order], None, frame_size=frame,
                                                     embed_batch=lambda batch, **kw: batch)
    host = {"items": host["items"], "ratings": host["ratings"], "sizes": host["sizes"],
            "table": torch.from_numpy(inp["table"]).to(dev)}
    n = host["items"].shape[0]
    s_dim, a_dim, h = C.dims(spec)
    masks = [torch.from_numpy(rng.integers(0, 2, size=(n, h)).astype(np.uint8)) for _ in range(8)]
    noise = torch.from_numpy((rng.standard_normal((n, a_dim)) * 0.2).astype(np.float32))
    update = recnn_b200.nn.ddpg_update if algo == "ddpg" else recnn_b200.nn.td3_update
    params = dict(C.DDPG_PARAMS if algo == "ddpg" else C.TD3_PARAMS)
    results = []
    for src in ("host", "feed"):
        nets = build_nets(spec, inp, dev)
        opts = build_optimizers("adam", nets, algo)
        losses = []
        for step in range(3):
            batch = dict(host) if src == "host" else dict(feed.batch(order))
            batch["dropout_masks"] = masks[:6 if algo == "ddpg" else 8]
            if algo == "td3":
                batch["noise"] = noise
            losses.append(update(batch, params, nets, opts, dev, {}, recnn_b200.utils.DummyWriter(), learn=True,
                                 step=step))
        results.append((losses, {k: dump_net(m) for k, m in nets.items()}))
    (l0, w0), (l1, w1) = results
    assert l0 == l1 and np.isfinite(l0[0]["policy"])
    for name in w0:
        for k in w0[name]:
            assert np.array_equal(w0[name][k], w1[name][k]), (name, k)
