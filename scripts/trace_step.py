"""Kernel timeline of live update steps (CUDA-graph replays) from torch.profiler / CUPTI.

    python scripts/trace_step.py [ddpg|td3] [out.json]

Writes, for a few consecutive non-policy steps of the bench workload, every kernel's (name, stream, start us, duration
us) relative to the step's first kernel -- the picture ncu's serialised launch list cannot give: which kernels overlap,
how long the gaps between dependent kernels are, where the critical path runs."""
from __future__ import annotations

import json
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    algo = sys.argv[1] if len(sys.argv) > 1 else "ddpg"
    out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/trace_%s.json" % algo
    dev = torch.device("cuda:0")
    b = bench.Bench(algo, 4096, dev, 0, 1, data_parallel=False)
    b.prime(False)
    agent = b.agent
    agent._step = 1
    for i in range(30):
        agent.update(b.batch(i % 16, False), learn=True)
        agent._step = 1 + (i % 8)            # stay on non-policy steps
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(6):
            b.flush.zero_()
            agent._step = 1 + i
            agent.update(b.batch(i, False), learn=True)
        torch.cuda.synchronize()
    tmp = out + ".chrome.json"
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    os.remove(tmp)
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ks.sort(key=lambda e: e["ts"])
    steps, cur = [], None
    for e in ks:
        name = re.sub(r"\(.*", "", e["name"])
        name = re.sub(r"recnn::|tc::|void ", "", name)
        if "zero_pad_columns" in name or (cur is None and "frame_gather" in name):
            cur = []
            steps.append(cur)
        if cur is not None:
            cur.append({"name": name[:60], "stream": e.get("args", {}).get("stream"), "ts": e["ts"], "dur": e["dur"]})
    res = []
    for st in steps:
        st = [k for k in st if "vectorized_elementwise" not in k["name"] and "fill" not in k["name"].lower()]
        if len(st) < 10:
            continue
        t0 = st[0]["ts"]
        last = max(k["ts"] + k["dur"] for k in st)
        res.append({"span_us": last - t0, "kernels": [dict(k, ts=round(k["ts"] - t0, 2)) for k in st]})
    json.dump(res, open(out, "w"), indent=0)
    for r in res[1:3]:
        print("step span %.1f us, %d kernels" % (r["span_us"], len(r["kernels"])))
        for k in r["kernels"]:
            print("  %8.1f %7.1f  s%-3s %s" % (k["ts"], k["dur"], k["stream"], k["name"]))


if __name__ == "__main__":
    main()
