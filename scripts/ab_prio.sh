#!/bin/bash
# A/B of the launch-priority experiment (RECNN_B200_PRIO = 0 / 1 / 2): bench value + live timeline
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
tag=${1:-prio}
for m in 0 1 2; do
  RECNN_B200_PRIO=$m timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-other-algo > $O/${tag}_bench_$m.json 2> $O/${tag}_bench_$m.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${tag}_bench_$m.json").read().strip().splitlines()[-1])
    print("PRIO=$m value %.1f (min %.1f max %.1f) e2e %.1f strong %.1f" % (d["value"], d["spread"]["min"], d["spread"]["max"], d["e2e"]["value"], d.get("strong", {}).get("updates_per_sec", 0)))
except Exception as e:
    print("PRIO=$m failed", e); print(open("$O/${tag}_bench_$m.err").read()[-1500:])
PY
  RECNN_B200_PRIO=$m timeout 120 python scripts/trace_step.py ddpg $O/${tag}_trace_$m.json > $O/${tag}_trace_$m.txt 2>&1
  grep "step span" $O/${tag}_trace_$m.txt | head -3
done
RECNN_B200_PRIO=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x -k "not data_parallel" > $O/${tag}_tests_1.log 2>&1
tail -2 $O/${tag}_tests_1.log
