#!/usr/bin/env python
"""Benchmark of the DDPG update hot path (BASELINE.json metric:
"DDPG update-steps/sec @ batch 4096 ...; embed-gather HBM GB/s").

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores
                                                           # (numpy port in oracle/; /root/reference is Python
                                                           # and does not exist on the GPU box)

One "step" = one DDPG update (ddpg_update) over one synthetic ML-20M-shaped minibatch:
26,744 items x 128-d table, frame_size 10, 4096 sample rows per GPU, policy step every 10th
step, Adam(lr=1e-5), dropout active (perf mode: on-device Philox).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ITEMS, DIM, FRAME, HIDDEN = 26744, 128, 10, 256
S_DIM = DIM * FRAME + FRAME
ROWS_PER_GPU = 4096
POLICY_STEP = 10
# algorithmic work per sample row
GATHER_BYTES_PER_ROW = 16604
DDPG_FLOP_NONPOLICY, DDPG_FLOP_POLICY = 5276160, 6526976
L1_FWD_FLOP_PER_ROW = 2 * S_DIM * HIDDEN           # the dominant kernel: layer-1 forward GEMM


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()          # exact PID, never by pattern
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
            sm = [float(r[0]) for r in rows]
            out["samples"] = len(sm)
            if sm:
                busy = [x for x in sm if x > 0.5 * max(sm)] or sm
                out["sm_mhz"] = float(np.median(busy))
                out["sm_max_mhz"] = float(rows[0][1])
                out["power_w_max"] = max(float(r[2]) for r in rows)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nm in enumerate(names):
                    if any(r[3 + i].strip().lower().startswith("active") for r in rows):
                        out["reasons"].append(nm)
        except Exception:
            pass
        return out


def synth_step_inputs(seed, n_rows, n_steps):
    """Per-step minibatches: items ~ U{0..n_items-1}, ratings ~ U{-4..5}, one pseudo-user."""
    rng = np.random.default_rng(seed)
    items = rng.integers(0, N_ITEMS, size=(n_steps, n_rows, FRAME + 1), dtype=np.int64)
    ratings = rng.integers(-4, 6, size=(n_steps, n_rows, FRAME + 1)).astype(np.float32)
    done = np.zeros((n_rows,), dtype=np.float32)
    done[-1] = 1.0
    return items, ratings, done


# =============================================================================== reference arm
def run_reference(args):
    """`--impl reference`: the reference algorithm on the host CPU (numpy port in oracle/; /root/reference is
    Python and is not on the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rows = ROWS_PER_GPU * args.gpus
    stepper = _OracleStepper(n_rows)
    limiter, threads = pick_blas_threads(stepper)
    ctx = limiter(limits=threads) if limiter is not None else None
    if ctx is not None:
        ctx.__enter__()
    try:
        for _ in range(args.warmup):
            stepper.run()
        t_total = sum(stepper.run() for _ in range(args.steps))
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)
    value = args.steps / t_total
    line = {
        "impl": "reference", "metric": "ddpg_update_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DDPG batch %d rows, 128-d embeddings, 26744 items, frame 10" % n_rows,
                   "rows_per_step": n_rows, "optimizer": "adam lr=1e-5", "policy_step": POLICY_STEP},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "host_cpus": os.cpu_count() or 1,
                         "kind": "port",
                         "sample": "%d full steps (gather + ddpg_update, numpy/OpenBLAS fp32, best-of thread count)" % args.steps},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


class _OracleStepper:
    """Reference algorithm (gather + ddpg_update) on the host: numpy port in oracle/."""

    def __init__(self, n_rows):
        from oracle import recnn_oracle as O
        from oracle import cases as C
        self.O, self.C, self.n = O, C, n_rows
        rng = np.random.default_rng(0)
        self.rng = rng
        self.table = rng.standard_normal((N_ITEMS, DIM), dtype=np.float32)
        self.nets = {"policy_net": O.make_actor(rng, S_DIM, DIM, HIDDEN, 6e-1),
                     "value_net": O.make_critic(rng, S_DIM, DIM, HIDDEN, 54e-2)}
        self.nets["target_policy_net"] = O.copy_net(self.nets["policy_net"])
        self.nets["target_value_net"] = O.copy_net(self.nets["value_net"])
        self.opts = {"policy_optimizer": O.make_optimizer("adam", lr=1e-5),
                     "value_optimizer": O.make_optimizer("adam", lr=1e-5)}
        self.sizes = np.asarray([n_rows + FRAME], dtype=np.int64)
        self.items, self.ratings, _ = synth_step_inputs(1, n_rows, 4)
        self.step = 0

    def run(self):
        """One step; returns its wall time (the dropout draw is not timed: RNGs differ per implementation)."""
        masks = self.O.synth_masks(self.rng, 6, self.n, HIDDEN)
        i = self.step % 4
        t0 = time.perf_counter()
        batch = self.O.frame_gather(self.table, self.items[i], self.ratings[i], self.sizes, FRAME)
        self.O.ddpg_update(batch, dict(self.C.DDPG_PARAMS), self.nets, self.opts, masks, self.step, learn=True)
        dt = time.perf_counter() - t0
        self.step += 1
        return dt


def pick_blas_threads(stepper):
    """One BLAS thread per hardware thread is far from OpenBLAS's best on a 100+ core box for GEMMs of this
    size; give the CPU arm the thread count that maximises ITS step rate (2 steps at each candidate)."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        return None, os.cpu_count() or 1
    ncpu = os.cpu_count() or 1
    best, best_t = ncpu, None
    for n in sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu}):
        with threadpool_limits(limits=n):
            stepper.run()
            t = min(stepper.run(), stepper.run())
        if best_t is None or t < best_t:
            best, best_t = n, t
    return threadpool_limits, best


def cpu_baseline_sample(seconds_budget=15.0):
    """Oracle port on the host cores, bounded: N=4096 DDPG steps until ~budget is used."""
    stepper = _OracleStepper(ROWS_PER_GPU)
    limiter, threads = pick_blas_threads(stepper)
    ctx = limiter(limits=threads) if limiter is not None else None
    if ctx is not None:
        ctx.__enter__()
    try:
        timed, t_total = 0, 0.0
        t_start = time.perf_counter()
        while True:
            t_total += stepper.run()
            timed += 1
            if (timed >= 10 and time.perf_counter() - t_start > seconds_budget) or timed >= 60:
                break
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)
    return {"value": timed / t_total, "unit": "steps/s", "cores": threads, "host_cpus": os.cpu_count() or 1,
            "kind": "port",
            "sample": "%d DDPG steps at 4096 rows (gather + update, numpy/OpenBLAS fp32, best-of thread count)" % timed}


# =============================================================================== native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    import recnn_b200
    from recnn_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    opts = {}
    for kv in args.opt:
        name, _, val = kv.partition("=")
        _lib.set_option(name, int(val))
        opts[name] = int(val)

    torch.manual_seed(1234)                       # same weights on every rank
    rng = np.random.default_rng(0)
    table = torch.from_numpy(rng.standard_normal((N_ITEMS, DIM), dtype=np.float32)).to(dev)
    actor = recnn_b200.nn.Actor(S_DIM, DIM, HIDDEN, 6e-1)
    critic = recnn_b200.nn.Critic(S_DIM, DIM, HIDDEN, 54e-2)
    agent = recnn_b200.nn.DDPG(actor, critic).to(dev)
    for k, net in (("policy_optimizer", "policy_net"), ("value_optimizer", "value_net")):
        agent.optimizers[k] = recnn_b200.optim.Adam(agent.nets[net].parameters(), lr=1e-5)
    if world > 1:
        recnn_b200.dist.enable_data_parallel(agent)

    n_rows = ROWS_PER_GPU
    total = args.steps + args.warmup
    n_distinct = min(total, 16)
    items_np, ratings_np, done_np = synth_step_inputs(100 + rank, n_rows, n_distinct)
    items_h = [torch.from_numpy(items_np[i]).pin_memory() for i in range(n_distinct)]
    ratings_h = [torch.from_numpy(ratings_np[i]).pin_memory() for i in range(n_distinct)]
    done_h = torch.from_numpy(done_np).pin_memory()
    items_d = [t.to(dev) for t in items_h]
    ratings_d = [t.to(dev) for t in ratings_h]
    done_d = done_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)    # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run_loop(host_inputs, steps, warmup, flush_l2=True):
        agent._step = 0
        eng_kernels0 = None
        times = []
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for it in range(warmup + steps):
            i = it % n_distinct
            if it == warmup:
                barrier()
                from recnn_b200.nn.update._engine import get_engine
                eng = get_engine(_lib.ALGO_DDPG, agent.nets, dev)
                eng_kernels0 = eng.kernels
                t_wall0 = time.perf_counter()
            if flush_l2:
                flush.zero_()
            batch = {"items": items_h[i] if host_inputs else items_d[i],
                     "ratings": ratings_h[i] if host_inputs else ratings_d[i],
                     "done": done_h if host_inputs else done_d, "table": table}
            if it >= warmup:
                ev0[it - warmup].record()
            loss = agent.update(batch, learn=True)          # H2D (if host) + fused step + D2H of the losses
            agent.step()
            if it >= warmup:
                ev1[it - warmup].record()
        barrier()
        wall = time.perf_counter() - t_wall0
        dev_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        eng = get_engine(_lib.ALGO_DDPG, agent.nets, dev)
        return float(t.item()), wall, eng.kernels - eng_kernels0, loss

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, wall_s, kernels, last_loss = run_loop(False, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else {}
    e2e_ms, e2e_wall, _, _ = run_loop(True, args.steps, max(3, args.warmup // 2))
    warm_ms, _, _, _ = run_loop(False, args.steps, 3, flush_l2=False)

    rows_global = n_rows * world
    # One unit = one 4096-row minibatch through the update step.  Data parallel is weak scaling: every rank
    # takes its own 4096 rows per step and the gradients are all-reduced, so a global step consumes `world`
    # units; `value` counts units/s over the whole job (= optimizer updates/s * n_gpus).
    global_steps_per_sec = args.steps / (dev_ms / 1e3)
    value = global_steps_per_sec * world
    e2e_value = args.steps / (e2e_ms / 1e3) * world
    line = None
    if rank == 0:
        L = _lib.lib()
        st = torch.cuda.current_stream(dev).cuda_stream

        def time_kernel(fn, iters=20):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize(dev)
                ts.append(a.elapsed_time(b))
            return float(np.mean(ts))

        # (1) the materialising gather kernel  (HBM bound)
        g_state = torch.empty(n_rows, S_DIM, device=dev)
        g_next = torch.empty(n_rows, S_DIM, device=dev)
        g_act = torch.empty(n_rows, DIM, device=dev)
        g_rew = torch.empty(n_rows, device=dev)
        gather_ms = time_kernel(lambda: _lib.check(L.recnn_frame_gather(
            table.data_ptr(), N_ITEMS, DIM, items_d[0].data_ptr(), ratings_d[0].data_ptr(), n_rows, FRAME,
            g_state.data_ptr(), g_next.data_ptr(), g_act.data_ptr(), g_rew.data_ptr(), None, st)))
        gather_gbs = n_rows * GATHER_BYTES_PER_ROW / (gather_ms * 1e-3) / 1e9
        # the same kernel on a batch whose output does not fit in L2 (16x the rows): its HBM-bound regime
        big = 16 * n_rows
        gb_items = torch.randint(0, N_ITEMS, (big, FRAME + 1), device=dev, dtype=torch.int64)
        gb_ratings = torch.rand(big, FRAME + 1, device=dev)
        gb_state = torch.empty(big, S_DIM, device=dev)
        gb_next = torch.empty(big, S_DIM, device=dev)
        gb_act = torch.empty(big, DIM, device=dev)
        gb_rew = torch.empty(big, device=dev)
        gather_big_ms = time_kernel(lambda: _lib.check(L.recnn_frame_gather(
            table.data_ptr(), N_ITEMS, DIM, gb_items.data_ptr(), gb_ratings.data_ptr(), big, FRAME,
            gb_state.data_ptr(), gb_next.data_ptr(), gb_act.data_ptr(), gb_rew.data_ptr(), None, st)), iters=10)
        gather_big_gbs = big * GATHER_BYTES_PER_ROW / (gather_big_ms * 1e-3) / 1e9
        del gb_state, gb_next, gb_act, gb_rew, gb_items, gb_ratings
        # (2) the dominant kernel of the step: layer-1 forward GEMM [4096,1290] x [1290,256]
        #     (same tcgen05 3xTF32 kernel and operand pitches as inside the step; plain-store epilogue)
        ld_s = (S_DIM + 3) // 4 * 4
        x_img = torch.randn(n_rows, ld_s, device=dev)
        w1 = agent.nets["policy_net"].linear1.weight            # strided view into the arena, pitch 1292
        h1 = torch.empty(n_rows, HIDDEN, device=dev)
        l1_ms = time_kernel(lambda: _lib.check(L.recnn_gemm_tf32x3(
            n_rows, HIDDEN, S_DIM, x_img.data_ptr(), ld_s, 0, w1.data_ptr(), w1.stride(0), 0,
            h1.data_ptr(), HIDDEN, 64, st)))
        l1_tflops = n_rows * L1_FWD_FLOP_PER_ROW / (l1_ms * 1e-3) / 1e12
        tf32_peak = peaks["bf16"] / 2.0           # dense TF32 = half the dense bf16 rate
        flop_step = rows_global * (DDPG_FLOP_POLICY + (POLICY_STEP - 1) * DDPG_FLOP_NONPOLICY) / POLICY_STEP
        feed_info = bench_device_feed(agent, table, dev, flush, time_kernel, args) if world == 1 else None
        cpu = cpu_baseline_sample() if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": "ddpg_update_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DDPG batch 4096 rows/GPU, 128-d embeddings, 26744 items, frame 10 (BASELINE configs[1])",
                       "rows_per_gpu": n_rows, "global_rows": rows_global, "parallelism": "dp%d" % world,
                       "unit_def": "4096-row minibatches through the update step per second, whole job "
                                   "(each data-parallel step consumes n_gpus of them; optimizer updates/s = value / n_gpus)",
                       "grad_allreduce": ("none" if world == 1 else
                                          "in-graph NVLink peer-memory kernels (recnn_comm_*)" if getattr(
                                              agent.nets["policy_net"], "_recnn_dp", (0, 0, None))[2] is not None
                                          else "NCCL between the step's phases"),
                       "optimizer": "adam lr=1e-5 (fused)", "policy_step": POLICY_STEP,
                       "dropout": "on (device Philox)", "l2": "flushed between timed steps (256 MB write)",
                       "inputs": "items/ratings/done resident in HBM; frames gathered on device inside the step",
                       "matmul": "tcgen05 3xTF32 (error-compensated, fp32-grade) with fp32 CUDA-core fallbacks for the 256->1 head"},
            "optimizer_updates_per_sec": global_steps_per_sec,
            "rows_per_sec": global_steps_per_sec * rows_global,
            "update_tflops": flop_step * global_steps_per_sec / 1e12,
            # (algorithmic FLOPs x 3 TF32 passes) / (t_step x TF32 peak x GPUs): the whole step, not one kernel
            "step_tensor_frac": 3.0 * flop_step * global_steps_per_sec / 1e12 / (tf32_peak * world),
            "value_warm_l2": args.steps / (warm_ms / 1e3) * world,
            "wall_s": wall_s,
            "e2e": {"value": e2e_value, "unit": "steps/s",
                    "h2d_bytes_per_step": int(n_rows * ((FRAME + 1) * 12 + 4)), "d2h_bytes_per_step": 16,
                    "what": "ddpg_update(batch of pinned host items/ratings/done) -> dict of python floats"},
            "gpu_launches": int(kernels),
            "roofline": {"bound": "tensor", "kernel": "layer-1 forward GEMM [4096x1290]x[1290x256] (tc_gemm_kernel, tcgen05 kind::tf32, 3 MMA passes/product)",
                         "achieved": 3.0 * l1_tflops, "algorithmic_fp32": l1_tflops, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": 3.0 * l1_tflops / tf32_peak,
                         "traffic": 22525184, "traffic_source": "dram__bytes_read+write per launch, profiles/r1b_ncu_tc_gemm_tile64_summary.csv (ncu --set full)", "peak_source": "%s bf16 %.0f TF/s / 2 (TF32 kind)" % (peaks["source"], peaks["bf16"]),
                         "ms": l1_ms},
            "roofline_gather": {"bound": "hbm", "kernel": "frame_gather_kernel", "achieved": gather_gbs,
                                "peak": peaks["hbm"], "unit": "GB/s", "frac": gather_gbs / peaks["hbm"],
                                "traffic": 12083712, "traffic_source": "dram__bytes_read+write per launch, profiles/README.md: the 44 MB of "
                                "output is absorbed by the 126 MB L2 inside the kernel, so DRAM traffic << algorithmic bytes", "peak_source": peaks["source"], "ms": gather_ms,
                                "bytes_per_launch": n_rows * GATHER_BYTES_PER_ROW,
                                "at_16x_rows": {"rows": big, "ms": gather_big_ms, "algorithmic_gbs": gather_big_gbs,
                                                "dram_gbs_est": (big * 10840 + N_ITEMS * DIM * 4) / (gather_big_ms * 1e-3) / 1e9,
                                                "dram_frac_est": (big * 10840 + N_ITEMS * DIM * 4) / (gather_big_ms * 1e-3) / 1e9 / peaks["hbm"],
                                                "note": "same kernel, 16x the rows: the 710 MB of output no longer fits in L2 and goes to HBM, "
                                                        "the 13.7 MB table is still served by L2, so DRAM traffic ~ output + table once"}},
            "clocks": clocks,
            "last_loss": last_loss,
        }
        if opts:
            line["config"]["options"] = opts
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if feed_info is not None:
            line["feed"] = feed_info
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


FEED_BYTES_PER_ROW = (FRAME + 1) * 12 * 2 + 4          # read ids+ratings, write ids+ratings, write done


def bench_device_feed(agent, table, dev, flush, time_kernel, args):
    """minibatches cut on the device out of resident user histories
    (recnn_b200.data.DeviceFrameFeed) instead of a DataLoader worker + H2D.  Synthetic "rolling users":
    2048 users x 138 interactions (128 windows each, 262,144 windows).  Reports the window-gather kernel
    alone and the update step fed by ``feed.sample(4096)`` (no host->device traffic at all)."""
    import torch
    from recnn_b200.data.feed import HistoryCSR, DeviceFrameFeed
    rng = np.random.default_rng(7)
    n_users, length = 2048, 138
    items = rng.integers(0, N_ITEMS, size=(n_users, length), dtype=np.int64)
    rates = rng.integers(-4, 6, size=(n_users, length)).astype(np.float64)
    feed = DeviceFrameFeed(HistoryCSR(np.arange(n_users), list(items), list(rates), FRAME), table, dev)
    n_rows = ROWS_PER_GPU
    w = torch.randint(0, feed.csr.n_windows, (n_rows,), device=dev)
    ids_ms = time_kernel(lambda: feed.windows(w))
    users32 = list(range(0, 32 * 8, 8))                    # 32 users x 128 windows = 4096 rows
    users_ms = time_kernel(lambda: feed.batch(users32))
    steps, warm = args.steps, max(3, args.warmup // 2)
    agent._step = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for it in range(warm + steps):
        flush.zero_()
        if it >= warm:
            ev[it - warm][0].record()
        agent.update(feed.sample(n_rows), learn=True)      # randint + window gather + fused step + loss read-back
        agent.step()
        if it >= warm:
            ev[it - warm][1].record()
    torch.cuda.synchronize(dev)
    ms = sum(a.elapsed_time(b) for a, b in ev)
    return {"what": "update step fed by DeviceFrameFeed.sample(4096): windows cut on the device from resident "
                    "histories (2048 users x 138 interactions), no host->device copies",
            "steps_per_sec": steps / (ms / 1e3), "ms_per_step": ms / steps, "h2d_bytes_per_step": 0,
            "window_gather_ids_ms": ids_ms, "window_gather_users_ms": users_ms,
            "window_gather_bytes_per_launch": n_rows * FEED_BYTES_PER_ROW,
            "note": "window_gather_*_ms include the output allocation and (users form) a 520-byte plan upload; "
                    "1.1 MB per launch: latency-bound, not bandwidth-bound"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="library A/B switch (recnn_debug_set_option), e.g. --opt presplit=1 --opt gather_variant=1")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (A/B runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 40:          # bounded: the CPU port does ~3-10 steps/s
            args.steps = 40
        args.warmup = min(args.warmup, 3)
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
