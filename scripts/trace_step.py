"""Timeline of the tensor-core kernels inside one (graph-replayed) DDPG step."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import recnn_b200
from recnn_b200 import _lib
from oracle import recnn_oracle as O
dev = torch.device("cuda:0")
h = ctypes.CDLL(_lib.lib_path())
h.recnn_debug_set_span.argtypes = [ctypes.c_void_p, ctypes.c_int]
h.recnn_debug_span_meta.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
rng = np.random.default_rng(0)
table, items, ratings, sizes = O.synth_frames(rng, 4096)
agent = recnn_b200.nn.DDPG(recnn_b200.nn.Actor(1290, 128, 256, 6e-1), recnn_b200.nn.Critic(1290, 128, 256, 54e-2)).to(dev)
batch = {"items": torch.from_numpy(items).to(dev), "ratings": torch.from_numpy(ratings).to(dev),
         "sizes": torch.from_numpy(sizes).to(dev), "table": torch.from_numpy(table).to(dev)}
span = torch.zeros(2 * 4096, dtype=torch.int64, device=dev)
h.recnn_debug_set_span(span.data_ptr(), 4096)
counts = []
for step in range(14):        # steps 0 (policy, eager), 1 (eager), 2 (capture) ... 10 (policy again)
    agent._step = step
    if step in (5, 13):
        span.fill_(0); span[0::2] = torch.iinfo(torch.int64).max
        torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); loss = agent.update(batch, learn=True); t1.record(); torch.cuda.synchronize()
    counts.append(h.recnn_debug_span_count())
    if step in (5,):
        ms = t0.elapsed_time(t1)
        s = span.cpu().numpy().reshape(-1, 2)
        # launches of the non-policy graph are the ones captured at step 2: slots [counts[1], counts[2])
        lo, hi = counts[1], counts[2]
        rows = []
        for i in range(lo, hi):
            m = (ctypes.c_longlong * 4)(); h.recnn_debug_span_meta(i, m)
            rows.append((s[i, 0], s[i, 1], m[0], m[1], m[2], m[3] & 0xffffffff, m[3] >> 32))
        rows = [r for r in rows if r[1] > 0]
        base = min(r[0] for r in rows)
        print("non-policy step: event time %.1f us, %d tensor-core launches" % (ms * 1000, len(rows)))
        prev_end = None
        for r in sorted(rows):
            cfg = r[2]
            print("  start %7.1f  dur %6.1f  end %7.1f  BN=%d A_MN=%d B_MN=%d EPI=%d  M=%d N=%d K=%d splits=%d" % (
                (r[0] - base) / 1e3, (r[1] - r[0]) / 1e3, (r[1] - base) / 1e3, cfg & 0xfff, (cfg >> 12) & 1, (cfg >> 13) & 1,
                cfg >> 16, r[3], r[4], r[5], r[6]))
        busy = sum(r[1] - r[0] for r in rows) / 1e3
        print("  sum of TC kernel durations %.1f us; span %.1f us" % (busy, (max(r[1] for r in rows) - base) / 1e3))
