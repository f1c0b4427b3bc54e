#!/bin/bash
# Multi-GPU session (gpurun --gpus N): data-parallel parity tests, the bench's N-GPU legs (dp_check, weak scaling,
# strong scaling) and the all-reduce latency microbenchmark.
#   usage: bash scripts/gpu_session_multi.sh <tag> "<world sizes to bench, e.g. 2 4 8>" "<pytest -k expression>"
cd "$(dirname "$0")/.." || exit 1
tag=${1:-r2m}
worlds=${2:-"2 4 8"}
kexpr=${3:-"data_parallel"}
mkdir -p gpurun_out
O=gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }
ng=$(python -c "import torch; print(torch.cuda.device_count())")
el "GPUs: $ng"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "$kexpr" > $O/${tag}_tests.log 2>&1
tail -4 $O/${tag}_tests.log
el "bench N=1"
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-other-algo > $O/${tag}_bench_n1.json 2> $O/${tag}_bench_n1.err
for n in $worlds; do
  [ $n -le $ng ] || continue
  el "bench N=$n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > $O/${tag}_bench_n$n.json 2> $O/${tag}_bench_n$n.err
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
    scripts/bench_comm.py 2> $O/${tag}_comm_n$n.err | tail -1 | tee $O/${tag}_comm_n$n.json
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("$O/${tag}_bench_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no result:", e); continue
    if n == 1: base = d["value"]
    st = d.get("strong", {})
    print("N=%d value %.1f (eff %.3f) ms/step %.4f  strong %.1f upd/s  dp_check %s" % (
        n, d["value"], d["value"] / (n * base) if base else 0, d["ms_per_step"], st.get("updates_per_sec", 0),
        d.get("dp_check", {}).get("status")))
PY
el "done"
