"""GPU parity: the CUDA path (public API -> C ABI -> sm_100a kernels) against
(a) the golden vectors generated from the unmodified reference and (b) the numpy
oracle run live on the same seeded inputs."""
import ctypes

import numpy as np
import pytest
import torch

import recnn_b200
from recnn_b200 import _lib
from oracle import cases as C
from oracle import recnn_oracle as O
from tests._golden import load_golden, compare_with_golden, run_oracle_case
from tests._cuda import run_cuda_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


# ----------------------------------------------------------------------------- gather
def test_gather_golden_bit_exact():
    g = load_golden("gather.npz")
    users = [{"items": g["user%d.items" % i], "rates": g["user%d.rates" % i],
              "sizes": len(g["user%d.items" % i]), "users": int(g["user%d.id" % i])} for i in range(3)]
    table = torch.from_numpy(g["table"]).to(DEV)
    out = recnn_b200.data.prepare_batch_static_size(users, table, frame_size=int(g["frame_size"]))
    for k in ("state", "next_state", "action", "reward", "done"):
        got = out[k].cpu().numpy()
        assert got.shape == g["out." + k].shape and got.dtype == np.float32
        assert np.array_equal(_bits(got), _bits(g["out." + k])), k
    assert np.array_equal(out["meta"]["sizes"].numpy(), g["out.sizes"])


@pytest.mark.parametrize("n_rows,n_items,dim,frame", [
    (1, 7, 128, 10), (33, 100, 128, 10), (4096, 26744, 128, 10), (257, 50, 16, 4),
    (64, 31, 7, 3),        # odd dim -> scalar path
    (100, 64, 256, 10),    # config-5 width
    (40, 20, 8, 33),       # frame+1 > 32 slots
])
def test_gather_vs_oracle(n_rows, n_items, dim, frame):
    rng = np.random.default_rng(n_rows * 31 + dim)
    table, items, ratings, sizes = O.synth_frames(rng, n_rows, n_items, dim, frame)
    ratings = (ratings + rng.standard_normal(ratings.shape).astype(np.float32)).astype(np.float32)
    want = O.frame_gather(table, items, ratings, sizes, frame)
    batch = {"items": torch.from_numpy(items), "ratings": torch.from_numpy(ratings),
             "sizes": torch.from_numpy(sizes), "users": torch.zeros(1, dtype=torch.int64)}
    got = recnn_b200.data.batch_tensor_embeddings(batch, torch.from_numpy(table).to(DEV), frame)
    for k in ("state", "next_state", "action", "reward", "done"):
        assert np.array_equal(_bits(got[k].cpu().numpy()), _bits(want[k])), k


@pytest.mark.parametrize("n_rows,n_items,dim,frame", [
    (1, 7, 128, 10), (4096, 26744, 128, 10), (4097, 500, 128, 10), (64, 31, 7, 3), (40, 20, 8, 33), (100, 64, 256, 10)])
def test_gather_bit_exact_and_out_of_range_ids(n_rows, n_items, dim, frame):
    """The gather kernel (one warp per row) on ragged shapes, and IndexError on an id == n_items."""
    rng = np.random.default_rng(n_rows + 7 * dim)
    table, items, ratings, sizes = O.synth_frames(rng, n_rows, n_items, dim, frame)
    want = O.frame_gather(table, items, ratings, sizes, frame)
    batch = {"items": torch.from_numpy(items), "ratings": torch.from_numpy(ratings),
             "sizes": torch.from_numpy(sizes), "users": torch.zeros(1, dtype=torch.int64)}
    got = recnn_b200.data.batch_tensor_embeddings(batch, torch.from_numpy(table).to(DEV), frame)
    bad = dict(batch, items=torch.from_numpy(np.where(items == items.max(), n_items, items)))
    with pytest.raises(IndexError):
        recnn_b200.data.batch_tensor_embeddings(bad, torch.from_numpy(table).to(DEV), frame)
    for k in ("state", "next_state", "action", "reward", "done"):
        assert np.array_equal(_bits(got[k].cpu().numpy()), _bits(want[k])), k


def test_gather_overlap_property_full_size():
    """FrameEnv-shaped rows: within a user next_state[i] == state[i+1]; done marks user ends."""
    rng = np.random.default_rng(5)
    users = [{"items": rng.integers(0, 26744, size=138, dtype=np.int64),
              "rates": rng.integers(-4, 6, size=138).astype(np.float64), "sizes": 138, "users": u}
             for u in range(32)]                                   # 32 * 128 = 4096 rows
    table = torch.from_numpy(rng.standard_normal((26744, 128), dtype=np.float32)).to(DEV)
    out = recnn_b200.data.prepare_batch_static_size(users, table, frame_size=10)
    s, s2, d = out["state"], out["next_state"], out["done"]
    assert s.shape == (4096, 1290)
    inner = torch.ones(4096, dtype=torch.bool, device=DEV)
    inner[127::128] = False
    assert torch.equal(s2[:-1][inner[:-1]], s[1:][inner[:-1]])
    assert torch.equal(d.nonzero().flatten().cpu(), torch.arange(127, 4096, 128))
    assert torch.equal(out["action"], table[torch.from_numpy(np.concatenate(
        [u["items"][10:] for u in users])).to(DEV)])


def test_gather_out_of_range_index_raises():
    table = torch.zeros(10, 16, device=DEV)
    batch = {"items": torch.full((4, 5), 10, dtype=torch.int64), "ratings": torch.zeros(4, 5),
             "sizes": torch.tensor([8]), "users": torch.tensor([0])}
    with pytest.raises(IndexError):
        recnn_b200.data.batch_tensor_embeddings(batch, table, 4)


# ----------------------------------------------------------------------------- forward
@pytest.mark.parametrize("n_rows", [1, 10, 300])
@pytest.mark.parametrize("train", [False, True])
def test_actor_critic_forward_vs_oracle(n_rows, train):
    rng = np.random.default_rng(11 + n_rows)
    pa = O.make_actor(rng, 1290, 128, 256, 6e-1)
    pc = O.make_critic(rng, 1290, 128, 256, 54e-2)
    s = rng.standard_normal((n_rows, 1290), dtype=np.float32)
    a = rng.standard_normal((n_rows, 128), dtype=np.float32)
    masks = O.synth_masks(rng, 2, n_rows, 256) if train else None
    from tests._cuda import load_net
    actor = load_net(recnn_b200.nn.Actor(1290, 128, 256), pa, DEV)
    critic = load_net(recnn_b200.nn.Critic(1290, 128, 256), pc, DEV)
    tm = [torch.from_numpy(m).to(DEV) for m in masks] if train else None
    actor.train(train), critic.train(train)
    got_a = actor(torch.from_numpy(s), masks=tm).cpu().numpy()
    got_q = critic(torch.from_numpy(s), torch.from_numpy(a), masks=tm).cpu().numpy()
    want_a, _ = O.actor_forward(pa, s, masks)
    want_q, _ = O.critic_forward(pc, s, a, masks)
    assert got_q.shape == (n_rows, 1)
    # fp32 GEMM vs fp32 GEMM: 1e-5 of the output scale (K=1290/1418 dot products)
    np.testing.assert_allclose(got_a, want_a, rtol=1e-5, atol=1e-5 * np.abs(want_a).max())
    np.testing.assert_allclose(got_q, want_q, rtol=1e-5, atol=1e-5 * np.abs(want_q).max())
    got_t = actor(torch.from_numpy(s), tanh=True, masks=tm).cpu().numpy()
    np.testing.assert_allclose(got_t, np.tanh(want_a), rtol=1e-5, atol=5e-6)   # tanhf vs libm tanh


def test_train_mode_forward_uses_dropout():
    actor = recnn_b200.nn.Actor(1290, 128, 256).to(DEV).train()
    s = torch.randn(64, 1290)
    assert not torch.equal(actor(s), actor(s))
    actor.eval()
    assert torch.equal(actor(s), actor(s))


# ----------------------------------------------------------------------------- small kernels
@pytest.mark.parametrize("tau", [1.0, 0.001, 1e-2])
def test_soft_update_vs_oracle(tau):
    rng = np.random.default_rng(3)
    p = O.make_actor(rng, 1290, 128, 256)
    t = O.make_actor(rng, 1290, 128, 256)
    from tests._cuda import load_net, dump_net
    net = load_net(recnn_b200.nn.Actor(1290, 128, 256), p, DEV)
    tgt = load_net(recnn_b200.nn.Actor(1290, 128, 256), t, DEV)
    recnn_b200.utils.soft_update(net, tgt, soft_tau=tau)
    want = O.copy_net(t)
    O.soft_update(p, want, tau)
    got = dump_net(tgt)
    for k in O.PARAM_ORDER:
        assert np.array_equal(_bits(got[k]), _bits(want[k])), k


@pytest.mark.parametrize("kind,kw", [("adam", dict(lr=1e-3)), ("adam", dict(lr=1e-5, weight_decay=1e-2)),
                                     ("sgd", dict(lr=1e-2)), ("sgd", dict(lr=1e-2, momentum=0.9, weight_decay=1e-3))])
def test_builtin_optimizer_vs_torch(kind, kw):
    """recnn_b200.optim.* against torch.optim.* on the same gradients (5 steps)."""
    torch.manual_seed(0)
    net = recnn_b200.nn.Critic(1290, 128, 256).to(DEV)
    ref = recnn_b200.nn.Critic(1290, 128, 256).to(DEV)
    ref.load_state_dict(net.state_dict())
    mine = (recnn_b200.optim.Adam if kind == "adam" else recnn_b200.optim.SGD)(net.parameters(), **kw).bind(net)
    theirs = (torch.optim.Adam if kind == "adam" else torch.optim.SGD)(ref.parameters(), **kw)
    from recnn_b200.nn.arena import grad_arena
    for it in range(5):
        grad_arena(net)                                   # p.grad are (strided) views of the arena
        for p, q in zip(net.parameters(), ref.parameters()):
            p.grad.copy_(torch.randn_like(p) * 0.01)
            q.grad = p.grad.clone()
        mine.step()
        theirs.step()
    for (n1, p1), (n2, p2) in zip(net.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(p1, p2, rtol=1e-6, atol=1e-8, msg=n1)
    assert mine.steps_taken() == 5


# ----------------------------------------------------------------------------- update steps
@pytest.mark.parametrize("case", ["tiny", "canon"])
@pytest.mark.parametrize("opt", ["adam", "sgd"])
@pytest.mark.parametrize("form", ["dense", "frames"])
def test_ddpg_vs_reference_golden(case, opt, form):
    gold = load_golden("ddpg_%s_%s.npz" % (case, opt))
    got = run_cuda_case(case, "ddpg", opt, form=form)
    compare_with_golden(got, gold)


@pytest.mark.parametrize("case", ["tiny", "canon"])
@pytest.mark.parametrize("opt", ["adam", "sgd"])
def test_td3_vs_reference_golden(case, opt):
    gold = load_golden("td3_%s_%s.npz" % (case, opt))
    got = run_cuda_case(case, "td3", opt, golden=gold, form="frames")
    compare_with_golden(got, gold)


@pytest.mark.parametrize("algo", ["ddpg", "td3"])
def test_external_torch_optimizer_path(algo):
    """torch.optim.Adam passed through the reference's `optimizer` dict (split-phase path)."""
    gold = load_golden("%s_canon_adam.npz" % algo)
    got = run_cuda_case("canon", algo, "adam", golden=gold, external=True)
    compare_with_golden(got, gold)


def test_cuda_vs_live_oracle_final_weights():
    """Every element of every tensor (not a sample) against the numpy oracle, tiny + canon."""
    for case in ("tiny", "canon"):
        want = run_oracle_case(case, "ddpg", "sgd")
        got = run_cuda_case(case, "ddpg", "sgd")
        for k in want:
            if k.startswith("final."):
                scale = np.abs(want[k]).max()
                np.testing.assert_allclose(got[k], want[k], rtol=1e-5, atol=1e-7 * scale + 1e-12, err_msg=k)


def test_quirks_on_device():
    """(1) the actor gradient left in .grad is sign-flipped and L1-normalised (ddpg.py:92);
    (2) TD3 never soft-updates its target policy (td3.py:136-141); (3) targets stay eval."""
    got = run_cuda_case("canon", "ddpg", "sgd")
    nets = got["_nets"]
    # after the last step (11, not a policy step) .grad still holds step 10's scaled gradient
    l1 = sum(p.grad.abs().sum().item() for p in nets["policy_net"].parameters())
    assert abs(l1 - 1.0) < 1e-4
    gold = load_golden("td3_canon_adam.npz")
    got3 = run_cuda_case("canon", "td3", "adam", golden=gold)
    inp = C.make_inputs(C.CASES["canon"], "td3")
    for k in O.PARAM_ORDER:
        assert np.array_equal(got3["final.target_policy_net." + k], inp["nets"]["target_policy_net"][k])
        assert not np.array_equal(got3["final.target_value_net1." + k], inp["nets"]["target_value_net1"][k])
    for name, m in got3["_nets"].items():
        assert m.training == ("target" not in name)


def test_external_optimizer_split_phases_meet_the_golden_bar():
    """External torch optimizers cut the step into several C calls (gradients complete -> optimizer.step() in
    Python -> next phase); the result must meet the same golden bar as the fused step."""
    gold = load_golden("ddpg_canon_adam.npz")
    got = run_cuda_case("canon", "ddpg", "adam", golden=gold, external=True)
    compare_with_golden(got, gold)


def test_graph_replay_equals_direct_launch(monkeypatch):
    from recnn_b200.nn.update import _engine
    a = run_cuda_case("canon", "ddpg", "adam", form="frames")
    monkeypatch.setattr(_engine, "_USE_GRAPHS", False)
    b = run_cuda_case("canon", "ddpg", "adam", form="frames")
    for k in a:
        if k.startswith("final.") or k.startswith("loss."):
            assert np.array_equal(a[k], b[k]), k


# ----------------------------------------------------------------------------- full size, perf mode
@pytest.mark.parametrize("algo", ["ddpg", "td3"])
def test_full_size_perf_mode_runs_and_learns(algo):
    """N=4096, D=128, 26,744 items, on-device Philox dropout/noise, Algo wrappers."""
    torch.manual_seed(1)
    rng = np.random.default_rng(9)
    table, items, ratings, sizes = O.synth_frames(rng, 4096)
    table_d = torch.from_numpy(table).to(DEV)
    actor = recnn_b200.nn.Actor(1290, 128, 256, 6e-1)
    if algo == "ddpg":
        agent = recnn_b200.nn.DDPG(actor, recnn_b200.nn.Critic(1290, 128, 256, 54e-2)).to(torch.device(DEV))
    else:
        agent = recnn_b200.nn.TD3(actor, recnn_b200.nn.Critic(1290, 128, 256, 54e-2),
                                  recnn_b200.nn.Critic(1290, 128, 256, 54e-2)).to(torch.device(DEV))
    for k in list(agent.optimizers):              # Adam with a bigger lr so 30 steps visibly reduce the loss
        net = agent.nets[k.replace("optimizer", "net")]     # (the default Ranger warms up for its first steps)
        agent.optimizers[k] = recnn_b200.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-2)
    batch = {"items": torch.from_numpy(items), "ratings": torch.from_numpy(ratings),
             "sizes": torch.from_numpy(sizes), "table": table_d}
    key = "value" if algo == "ddpg" else "value1"
    hist = []
    for _ in range(30):
        loss = agent.update(batch, learn=True)
        agent.step()
        assert all(np.isfinite(v) for v in loss.values())
        hist.append(loss[key])
    assert np.mean(hist[-5:]) < 0.7 * np.mean(hist[:5]), hist
    for name, net in agent.nets.items():
        assert net.training == ("target" not in name)
        assert all(torch.isfinite(p).all() for p in net.parameters())


def _spec(n_rows, **kw):
    return dict(C.FULL_SPEC, n_rows=n_rows, **kw)


@pytest.mark.parametrize("algo,spec", [
    ("ddpg", C.FULL_SPEC), ("td3", C.FULL_SPEC),
    # ragged shapes: a single row, one row past a 128-row tile, a row count that is no multiple of the
    # GEMM / split-K granules, and a narrow net (hidden 64, 32-d embeddings, frame 3)
    ("ddpg", _spec(1, n_items=500)), ("td3", _spec(1, n_items=500)),
    ("ddpg", _spec(129, n_items=2000)), ("td3", _spec(1000, n_items=5000)),
    ("ddpg", _spec(333, n_items=700, dim=32, frame=3, hidden=64)),
], ids=["ddpg-4096", "td3-4096", "ddpg-1row", "td3-1row", "ddpg-129", "td3-1000", "ddpg-narrow-333"])
def test_full_size_parity_vs_live_oracle(algo, spec):
    """BASELINE configs[1] (DDPG) / configs[2] (TD3) at FULL size -- 4096 rows, 26,744 x 128 table, frame
    10, H=256 -- and ragged row counts / a narrow net: three parity-mode steps (explicit dropout masks /
    TD3 noise, policy step at 0), frames form, SGD(1e-3), against the numpy oracle run live on the same inputs.

    Losses: the north-star 1e-5.  Weights: at this size a few of the ~6M ReLU pre-activations per step
    land within fp32 rounding error of 0, where two correct implementations may gate differently (the
    loss is continuous there, the gradient is not); one such flip moves ONE row of a first-layer weight
    gradient by ~1/32 of its norm.  So the bar on the weight CHANGES since init is: relative L2 error of
    every tensor's change <= 2e-2, and >= 90% of the elements of every tensor within 2e-3 of the
    tensor's largest change (the golden bar; it holds for every row without a flipped gate)."""
    want = run_oracle_case(spec, algo, "sgd")
    got = run_cuda_case(spec, algo, "sgd", form="frames")
    inp = C.make_inputs(spec, algo)
    for k in (k for k in want if k.startswith("loss.")):
        err = np.max(np.abs(got[k] - want[k]) / (np.abs(want[k]) + 0.1))
        assert err <= 1e-5, (k, err, got[k], want[k])
    stats = {}
    for k in (k for k in want if k.startswith("final.")):
        _, name, tensor = k.split(".")
        init = inp["nets"][name][tensor].astype(np.float64)
        d_want, d_got = want[k].astype(np.float64) - init, got[k].astype(np.float64) - init
        scale = np.max(np.abs(d_want))
        if scale == 0.0:                       # e.g. TD3's target policy is never updated (td3.py:136-141)
            assert np.array_equal(got[k], want[k]), k
            continue
        # differences below 2 ulp of the tensor's largest weight are fp32 rounding of the stored weight,
        # not signal (the actor's L1-normalised SGD update and the Polyak targets move by a few ulp only)
        ulp2 = 2.0 * 1.1920929e-07 * np.max(np.abs(want[k]))
        excess = np.maximum(np.abs(d_got - d_want) - ulp2, 0.0)
        l2 = np.linalg.norm(excess) / np.linalg.norm(d_want)
        ok = excess <= 2e-3 * scale
        stats[k] = (float(l2), float(ok.mean()))
        assert l2 <= 2e-2 and ok.mean() >= 0.90, (k, stats[k])
    assert len(stats) >= 12


def test_perf_mode_dropout_statistics():
    """Philox keep-rate is ~0.5 and masks differ between steps and layers."""
    torch.manual_seed(2)
    actor = recnn_b200.nn.Actor(1290, 128, 256).to(DEV).train()
    s = torch.randn(512, 1290, device=DEV)
    # forward() in train mode draws torch masks; the step's Philox stream is covered above.  Here:
    out = actor(s)
    assert torch.isfinite(out).all()


def test_learn_false_fills_debug_and_does_not_train():
    got_nets = run_cuda_case("tiny", "ddpg", "adam")["_nets"]
    before = {k: [p.clone() for p in m.parameters()] for k, m in got_nets.items()}
    spec = C.CASES["tiny"]
    inp = C.make_inputs(spec, "ddpg")
    ref = O.frame_gather(inp["table"], inp["items"], inp["ratings"], inp["sizes"], spec["frame"])
    batch = {k: torch.from_numpy(v) for k, v in ref.items()}
    debug = {}
    opts = {"policy_optimizer": None, "value_optimizer": None}
    loss = recnn_b200.nn.ddpg_update(batch, dict(C.DDPG_PARAMS), got_nets, opts, torch.device(DEV), debug,
                                     recnn_b200.utils.DummyWriter(), learn=False, step=0)
    assert set(loss) == {"value", "policy", "step"} and loss["value"] > 0
    assert debug["next_action"].shape == (spec["n_rows"], spec["dim"])
    assert debug["gen_action"].shape == (spec["n_rows"], spec["dim"])
    for k, m in got_nets.items():
        for p, q in zip(m.parameters(), before[k]):
            assert torch.equal(p, q)
    with pytest.raises(TypeError):
        recnn_b200.nn.ddpg_update(batch, dict(C.DDPG_PARAMS), got_nets, opts, torch.device(DEV), None,
                                  recnn_b200.utils.DummyWriter(), learn=False, step=0)


def test_cpu_device_is_rejected_loudly():
    spec = C.CASES["tiny"]
    s_dim, a_dim, h = C.dims(spec)
    nets = {"policy_net": recnn_b200.nn.Actor(s_dim, a_dim, h), "target_policy_net": recnn_b200.nn.Actor(s_dim, a_dim, h),
            "value_net": recnn_b200.nn.Critic(s_dim, a_dim, h), "target_value_net": recnn_b200.nn.Critic(s_dim, a_dim, h)}
    with pytest.raises(_lib.RecnnError):
        recnn_b200.nn.ddpg_update({}, dict(C.DDPG_PARAMS), nets, {}, torch.device("cpu"), {}, learn=True, step=0)


# ----------------------------------------------------------------------------- data parallel (2 GPUs)
def _dp_worker(rank, world, port, algo, transport, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RECNN_B200_COMM"] = transport
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        gold = load_golden("%s_canon_adam.npz" % algo)
        got = run_cuda_case("canon", algo, "adam", golden=gold, form="frames", device="cuda:%d" % rank,
                            shard=(rank, world))
        nets = got.pop("_nets")
        comm = nets["policy_net"].__dict__["_recnn_dp"][2]
        got["_peer_comm"] = np.asarray(comm is not None)
        if comm is not None:
            # the collective on its own: odd sizes (scalar path), the arena size (float4 path), repeated calls
            gen = torch.Generator(device="cuda:%d" % rank).manual_seed(100 + rank)
            for n in (3, 1001, 429828, 3, 4096):
                x = torch.randn(n, device="cuda:%d" % rank, generator=gen)
                want = x.clone()
                dist.all_reduce(want)
                comm.all_reduce(x)
                torch.cuda.synchronize()
                if world == 2:     # a + b has one possible rounding; more ranks: NCCL's tree order differs from rank order
                    assert torch.equal(x, want), "peer all-reduce != NCCL all-reduce (n=%d)" % n
                else:
                    assert torch.allclose(x, want, rtol=1e-5, atol=1e-5), "peer all-reduce != NCCL all-reduce (n=%d)" % n
                # identical bits on every rank
                x0 = x.clone()
                dist.broadcast(x0, src=0)
                assert torch.equal(x, x0), "peer all-reduce differs between ranks (n=%d)" % n
        q.put((rank, {k: v for k, v in got.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("algo", ["ddpg", "td3"])
def test_data_parallel_equals_reference(algo, transport, world):
    """Rows sharded over `world` ranks + gradient all-reduce == the single-process reference (golden),
    and the replicas stay bit-identical.  transport: the in-graph NVLink peer-memory all-reduce
    (recnn_comm_*: two-shot, fused with the optimizer and the loss sums) or NCCL calls between the phases.
    The canonical case has 32 rows: 16 / 8 / 4 rows per rank."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, algo, transport, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gold = load_golden("%s_canon_adam.npz" % algo)
    for rank in range(world):
        assert bool(res[rank].pop("_peer_comm")) == (transport == "peer"), "wrong all-reduce transport was used"
        compare_with_golden(res[rank], gold, check_grads=False)
    for rank in range(1, world):
        for k in res[0]:
            if k.startswith("final."):
                assert np.array_equal(res[0][k], res[rank][k]), "replicas diverged: rank %d %s" % (rank, k)


# ----------------------------------------------------------------------------- full size, tight bar
@pytest.mark.parametrize("algo", ["ddpg", "td3"])
def test_full_size_tight_parity_without_ambiguous_gates(algo):
    """BASELINE configs[1] / [2] at FULL size (4096 rows, 26,744 x 128 table) held to the GOLDEN bar on EVERY weight.

    test_full_size_parity_vs_live_oracle has to tolerate the few ReLU gates per step that land within fp32
    rounding error of zero.  Here those gates are removed from the problem instead of from the bar: the oracle
    logs every kept unit with |pre-activation| <= 2e-5 (20x the difference between fp32 GEMM implementations),
    the replayed dropout masks drop exactly those units (tests/_golden.py:neutralise_ambiguous_gates), and the
    same masks go to the CUDA path.  With no ambiguous gate left, the north-star bar must hold for 100% of the
    elements: losses 1e-5; every weight 1e-5 relative (floor: 1e-7 of the tensor's largest weight); every weight
    CHANGE within 2e-3 of the tensor's largest change.  The number of units dropped (a few hundred out of
    ~19 million gates) is asserted to be small, which is the 'handful of flips' claim made measurable."""
    from tests._golden import neutralise_ambiguous_gates
    spec = C.FULL_SPEC
    inp, dropped, rounds = neutralise_ambiguous_gates(spec, algo, "sgd")
    gates = spec["n_rows"] * spec["hidden"] * 2 * (3 if algo == "ddpg" else 4) * spec["steps"]
    assert dropped <= 2e-4 * gates, (dropped, gates, rounds)
    want = run_oracle_case(spec, algo, "sgd", inp=inp)
    got = run_cuda_case(spec, algo, "sgd", form="frames", inp=inp)
    for k in (k for k in want if k.startswith("loss.")):
        err = np.max(np.abs(got[k] - want[k]) / (np.abs(want[k]) + 0.1))
        assert err <= 1e-5, (k, err, got[k], want[k])
    checked = 0
    for k in (k for k in want if k.startswith("final.")):
        _, name, tensor = k.split(".")
        init = inp["nets"][name][tensor].astype(np.float64)
        w_want, w_got = want[k].astype(np.float64), got[k].astype(np.float64)
        wmax = np.max(np.abs(w_want))
        rel = np.max(np.abs(w_got - w_want) / (np.abs(w_want) + 1e-2 * wmax))
        assert rel <= 1e-5, (k, rel)
        d_want, d_got = w_want - init, w_got - init
        scale = np.max(np.abs(d_want))
        if scale == 0.0:
            assert np.array_equal(got[k], want[k]), k
            continue
        ulp2 = 2.0 * 1.1920929e-07 * wmax
        excess = np.maximum(np.abs(d_got - d_want) - ulp2, 0.0)
        assert np.max(excess) <= 2e-3 * scale, (k, float(np.max(excess) / scale), dropped)
        checked += 1
    assert checked >= 12


# ----------------------------------------------------------------------------- value_update on its own
@pytest.mark.parametrize("case", ["tiny", "canon"])
@pytest.mark.parametrize("opt_kind", ["sgd", "adam"])
def test_value_update_standalone_vs_oracle(case, opt_kind):
    """recnn.nn.update.value_update (misc.py:10-55) as a public entry point: three critic-only steps against
    O.value_update on the same inputs -- loss tensor, the critic's weights, and nothing else moves."""
    spec = C.CASES[case]
    inp = C.make_inputs(spec, "ddpg")
    dev = torch.device(DEV)
    from tests._cuda import build_nets, build_optimizers, dump_net
    nets = build_nets(spec, inp, dev)
    opts = build_optimizers(opt_kind, nets, "ddpg")
    o_nets = {k: O.copy_net(v) for k, v in inp["nets"].items()}
    from tests._golden import oracle_optimizers
    o_opts = oracle_optimizers(opt_kind, "ddpg")
    ref = O.frame_gather(inp["table"], inp["items"], inp["ratings"], inp["sizes"], spec["frame"])
    params = dict(C.DDPG_PARAMS)
    before = {k: dump_net(m) for k, m in nets.items()}
    for step in range(3):
        masks = inp["masks"][step]
        want_loss, _ = O.value_update(ref, params, o_nets, o_opts, masks[0:2], learn=True)
        batch = {k: torch.from_numpy(v) for k, v in ref.items()}
        batch["dropout_masks"] = [torch.from_numpy(m) for m in masks]
        got_loss = recnn_b200.nn.update.value_update(batch, params, nets, opts, dev, {}, learn=True, step=step)
        assert torch.is_tensor(got_loss) and got_loss.dim() == 0
        assert abs(float(got_loss) - float(want_loss)) <= 1e-5 * (abs(float(want_loss)) + 0.1)
    for name, m in nets.items():
        after = dump_net(m)
        for t in O.PARAM_ORDER:
            if name != "value_net":
                assert np.array_equal(after[t], before[name][t]), (name, t)      # only the critic is stepped
                continue
            w = o_nets[name][t].astype(np.float64)
            wmax = np.max(np.abs(w))
            assert np.max(np.abs(after[t] - w) / (np.abs(w) + 1e-2 * wmax)) <= 1e-5, t
            d_want = w - before[name][t]
            scale = np.max(np.abs(d_want))
            assert scale > 0
            ulp2 = 2.0 * 1.1920929e-07 * wmax
            assert np.max(np.maximum(np.abs((after[t] - before[name][t]) - d_want) - ulp2, 0)) <= 2e-3 * scale, t


def test_value_update_learn_false_fills_debug():
    spec = C.CASES["tiny"]
    inp = C.make_inputs(spec, "ddpg")
    from tests._cuda import build_nets
    nets = build_nets(spec, inp, torch.device(DEV))
    ref = O.frame_gather(inp["table"], inp["items"], inp["ratings"], inp["sizes"], spec["frame"])
    batch = {k: torch.from_numpy(v) for k, v in ref.items()}
    debug = {}
    loss = recnn_b200.nn.update.value_update(batch, dict(C.DDPG_PARAMS), nets, {"value_optimizer": None},
                                             torch.device(DEV), debug, learn=False)
    want, dbg = O.value_update(ref, dict(C.DDPG_PARAMS), {k: O.copy_net(v) for k, v in inp["nets"].items()},
                               {}, None, learn=False)
    # eval-mode target nets; the online critic is in train mode but learn=False gives it no masks here:
    assert np.isfinite(float(loss))
    np.testing.assert_allclose(debug["next_action"].cpu().numpy(), dbg["next_action"], rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------- error reporting / optimizer plumbing
def test_in_step_gather_reports_out_of_range_ids():
    """an item id outside [0, n_items) raises IndexError like batch_tensor_embeddings does
    (the in-step gather used to clamp silently)."""
    torch.manual_seed(0)
    agent = recnn_b200.nn.DDPG(recnn_b200.nn.Actor(1290, 128, 256, 6e-1),
                               recnn_b200.nn.Critic(1290, 128, 256, 54e-2)).to(torch.device(DEV))
    rng = np.random.default_rng(1)
    table, items, ratings, sizes = O.synth_frames(rng, 64, 300)
    batch = {"items": torch.from_numpy(items), "ratings": torch.from_numpy(ratings),
             "sizes": torch.from_numpy(sizes), "table": torch.from_numpy(table).to(DEV)}
    agent.update(batch, learn=True)                      # fine
    bad = items.copy()
    bad[5, 3] = 300                                      # == n_items
    batch["items"] = torch.from_numpy(bad)
    with pytest.raises(IndexError):
        agent.update(batch, learn=True)
    bad[5, 3] = -1
    batch["items"] = torch.from_numpy(bad)
    with pytest.raises(IndexError):
        agent.update(batch, learn=True)
    batch["items"] = torch.from_numpy(items)
    assert np.isfinite(agent.update(batch, learn=True)["value"])      # the flag does not stick


def test_online_nets_must_share_train_mode():
    agent = recnn_b200.nn.DDPG(recnn_b200.nn.Actor(1290, 128, 256), recnn_b200.nn.Critic(1290, 128, 256)).to(torch.device(DEV))
    rng = np.random.default_rng(1)
    table, items, ratings, sizes = O.synth_frames(rng, 16, 300)
    batch = {"items": torch.from_numpy(items), "ratings": torch.from_numpy(ratings),
             "sizes": torch.from_numpy(sizes), "table": torch.from_numpy(table).to(DEV)}
    agent.nets["value_net"].eval()
    with pytest.raises(ValueError):
        agent.update(batch, learn=True)


def test_builtin_policy_with_external_value_optimizer_steps_each_once():
    """replacing only value_optimizer by a torch optimizer used to step the policy twice."""
    spec = C.CASES["tiny"]
    inp = C.make_inputs(spec, "ddpg")
    from tests._cuda import build_nets
    dev = torch.device(DEV)
    ref = O.frame_gather(inp["table"], inp["items"], inp["ratings"], inp["sizes"], spec["frame"])
    results = []
    for mixed in (False, True):
        nets = build_nets(spec, inp, dev)
        opts = {"policy_optimizer": recnn_b200.optim.SGD(nets["policy_net"].parameters(), lr=1e-3),
                "value_optimizer": (torch.optim.SGD(nets["value_net"].parameters(), lr=1e-3) if mixed else
                                    recnn_b200.optim.SGD(nets["value_net"].parameters(), lr=1e-3))}
        for step in range(2):
            batch = {k: torch.from_numpy(v) for k, v in ref.items()}
            batch["dropout_masks"] = [torch.from_numpy(m) for m in inp["masks"][step]]
            recnn_b200.nn.ddpg_update(batch, dict(C.DDPG_PARAMS), nets, opts, dev, {}, learn=True, step=step * 10)
        assert opts["policy_optimizer"].steps_taken() == 2
        results.append([p.detach().cpu().clone() for p in nets["policy_net"].parameters()])
    for p, q in zip(*results):
        assert torch.allclose(p, q, rtol=1e-6, atol=1e-8)


def test_optimizer_state_dict_round_trip_continues_bias_correction():
    spec = C.CASES["tiny"]
    inp = C.make_inputs(spec, "ddpg")
    from tests._cuda import build_nets, build_optimizers
    dev = torch.device(DEV)
    ref = O.frame_gather(inp["table"], inp["items"], inp["ratings"], inp["sizes"], spec["frame"])

    def run(steps, nets, opts, first=0):
        for step in range(first, first + steps):
            batch = {k: torch.from_numpy(v) for k, v in ref.items()}
            batch["dropout_masks"] = [torch.from_numpy(m) for m in inp["masks"][step]]
            recnn_b200.nn.ddpg_update(batch, dict(C.DDPG_PARAMS), nets, opts, dev, {}, learn=True, step=step)

    nets_a = build_nets(spec, inp, dev)
    opts_a = build_optimizers("adam", nets_a, "ddpg")
    run(4, nets_a, opts_a)
    # checkpoint after 2 steps, resume in fresh objects
    nets_b = build_nets(spec, inp, dev)
    opts_b = build_optimizers("adam", nets_b, "ddpg")
    run(2, nets_b, opts_b)
    sd_nets = {k: m.state_dict() for k, m in nets_b.items()}
    sd_opts = {k: o.state_dict() for k, o in opts_b.items()}
    assert sd_opts["value_optimizer"]["recnn_arenas"]["t"] == 2
    nets_c = build_nets(spec, inp, dev)
    for k, m in nets_c.items():
        m.load_state_dict(sd_nets[k])
    opts_c = build_optimizers("adam", nets_c, "ddpg")
    for k, o in opts_c.items():
        o.load_state_dict(sd_opts[k])
    run(2, nets_c, opts_c, first=2)
    for k in nets_a:
        for p, q in zip(nets_a[k].parameters(), nets_c[k].parameters()):
            assert torch.equal(p, q), k


# ----------------------------------------------------------------------------- Ranger (the reference's default optimizer)
@pytest.mark.parametrize("algo", ["ddpg", "td3"])
@pytest.mark.parametrize("case", ["tiny", "canon"])
def test_ranger_step_vs_oracle(case, algo):
    """recnn.nn.DDPG / TD3 build torch_optimizer.Ranger by default (algo.py:84-89).  The fused RANGER kind against
    the oracle's restatement (whose RAdam half is pinned against torch.optim.RAdam and whose Lookahead half against
    its definition, tests/test_oracle_golden.py) over 12 steps: rectification switch at step 6, Lookahead at 6 / 12."""
    want = run_oracle_case(case, algo, "ranger")
    got = run_cuda_case(case, algo, "ranger", form="frames")
    inp = C.make_inputs(C.CASES[case], algo)
    for k in (k for k in want if k.startswith("loss.")):
        err = np.max(np.abs(got[k] - want[k]) / (np.abs(want[k]) + 0.1))
        assert err <= 1e-5, (k, err)
    for k in (k for k in want if k.startswith("final.")):
        _, name, tensor = k.split(".")
        w = want[k].astype(np.float64)
        wmax = np.max(np.abs(w))
        assert np.max(np.abs(got[k] - w) / (np.abs(w) + 1e-2 * wmax)) <= 1e-5, k
        d_want = w - inp["nets"][name][tensor]
        scale = np.max(np.abs(d_want))
        if scale == 0:
            assert np.array_equal(got[k], want[k]), k
            continue
        ulp2 = 2.0 * 1.1920929e-07 * wmax
        excess = np.maximum(np.abs((got[k] - inp["nets"][name][tensor]) - d_want) - ulp2, 0)
        assert np.max(excess) <= 2e-3 * scale, (k, float(np.max(excess) / scale))


def test_default_optimizers_are_ranger_like_the_reference():
    agent = recnn_b200.nn.DDPG(recnn_b200.nn.Actor(44, 8, 32), recnn_b200.nn.Critic(44, 8, 32))
    for o in agent.optimizers.values():
        assert isinstance(o, recnn_b200.optim.Ranger)
        g = o.param_groups[0]
        assert (g["lr"], g["weight_decay"], g["k"], g["alpha"], g["betas"], g["eps"]) == (1e-5, 1e-2, 6, 0.5, (0.95, 0.999), 1e-5)
