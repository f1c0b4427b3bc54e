#!/bin/bash
# First GPU session of round 2: the LEAN GEMM kernels were written after round 1's GPU budget was spent.
#   1. their correctness tests (gated by RECNN_TEST_EXPERIMENTAL=1), run in a subprocess with a hard timeout:
#      every mbarrier wait in the kernels is bounded (trap after ~2 s), so a pipeline bug fails, it does not hang;
#   2. A/B timing of single GEMMs (default / presplit / 16 workers / lo2 / lean / lean+16 workers);
#   3. the step under RECNN_B200_LEAN=1 (+ WORKERS16=1), short benches;
#   4. if a lean configuration is correct and faster: full GPU suite + full bench under it.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
t0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - t0 ))s] $*"; }

el "1. experimental kernel tests"
RECNN_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -q -k "lean" \
  --maxfail=5 --tb=short > $O/t_lean.log 2>&1
RECNN_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -k "sixteen_byte or column_sums or one_wave" --tb=short > $O/t_gather16.log 2>&1
tail -2 $O/t_gather16.log
tail -5 $O/t_lean.log

el "2. A/B timing of single GEMMs"
timeout 120 python scripts/ab_gemm_variants.py > $O/ab_gemm_r2.json 2> $O/ab_gemm_r2.err
python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/ab_gemm_r2.json"))["ab_presplit"]:
        print(r["gemm"][:16], "tile", r["tile_n"], " ".join("%s=%.1f" % (k[:-10], v) for k, v in r.items() if k.endswith("_us_median")),
              "lean identical:", r.get("lean_bit_identical"), r.get("lean_w16_bit_identical"), r.get("lean_error", ""))
except Exception as e:
    print("ab failed:", e)
PY

el "3. step A/B"
timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_r2_default.json 2> $O/bench_r2_default.err
timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --opt gather_variant=2 > $O/bench_r2_gather16.json 2> $O/bench_r2_gather16.err
timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --opt tail=1 > $O/bench_r2_tail.json 2> $O/bench_r2_tail.err
timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --opt dwsplit=1 > $O/bench_r2_dwsplit.json 2> $O/bench_r2_dwsplit.err
timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --opt padzero=1 > $O/bench_r2_padzero.json 2> $O/bench_r2_padzero.err
python -c "
import json
for f in ('default','gather16','tail','dwsplit','padzero'):
    d=json.load(open('gpurun_out/bench_r2_%s.json'%f)); print(f, round(d['value'],1))"
i=0
for opts in "--opt lean=1" "--opt lean=1 --opt workers16=1" "--opt lean=1 --opt workers16=1 --opt bn64=1" \
            "--opt lean=1 --opt presplit=1" "--opt lean=1 --opt presplit=1 --opt workers16=1" \
            "--opt lean=1 --opt pdl=1" "--opt lean=1 --opt pdl=1 --opt workers16=1" \
            "--opt lean=1 --opt presplit=1 --opt pdl=1 --opt tail=1 --opt dwsplit=1 --opt padzero=1 --opt gather_variant=2" \
            "--opt lean=1 --opt presplit=1 --opt pdl=1 --opt tail=1 --opt dwsplit=1 --opt padzero=1 --opt gather_variant=2 --opt workers16=1"; do
  i=$((i+1))
  timeout 100 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $opts > $O/bench_r2_v$i.json 2> $O/bench_r2_v$i.err
done
python - <<'PY' > gpurun_out/best_env_r2.sh
import json, sys
log = open("gpurun_out/t_lean.log").read()
ok = (" passed" in log) and ("failed" not in log) and ("error" not in log.lower())
log2 = open("gpurun_out/t_gather16.log").read()
ok2 = (" passed" in log2) and ("failed" not in log2) and ("error" not in log2.lower())
sys.stderr.write("lean kernels correct: %s; gather16 / tail / dwsplit / padzero correct: %s (variants 8, 9 need both)\n" % (ok, ok2))
try:
    base = json.load(open("gpurun_out/bench_r2_default.json"))["value"]
except Exception:
    base = 2300.0
envs = {1: "RECNN_B200_LEAN=1", 2: "RECNN_B200_LEAN=1 RECNN_B200_WORKERS16=1",
        3: "RECNN_B200_LEAN=1 RECNN_B200_WORKERS16=1 RECNN_B200_BN64=1",
        4: "RECNN_B200_LEAN=1 RECNN_B200_PRESPLIT=1", 5: "RECNN_B200_LEAN=1 RECNN_B200_PRESPLIT=1 RECNN_B200_WORKERS16=1",
        6: "RECNN_B200_LEAN=1 RECNN_B200_PDL=1", 7: "RECNN_B200_LEAN=1 RECNN_B200_PDL=1 RECNN_B200_WORKERS16=1",
        8: "RECNN_B200_LEAN=1 RECNN_B200_PRESPLIT=1 RECNN_B200_PDL=1 RECNN_B200_TAIL=1 RECNN_B200_DWSPLIT=1 RECNN_B200_PADZERO=1 RECNN_B200_GATHER=2",
        9: "RECNN_B200_LEAN=1 RECNN_B200_PRESPLIT=1 RECNN_B200_PDL=1 RECNN_B200_TAIL=1 RECNN_B200_DWSPLIT=1 RECNN_B200_PADZERO=1 RECNN_B200_GATHER=2 RECNN_B200_WORKERS16=1"}
best, best_v = "", base * 1.02
for i in ((1, 2, 3, 4, 5, 6, 7, 8, 9) if ok else ()):
    if i >= 8 and not ok2:
        continue
    try:
        d = json.load(open("gpurun_out/bench_r2_v%d.json" % i))
        sys.stderr.write("variant %d (%s): %.1f steps/s, L1 gemm %.2f us\n" % (i, envs[i], d["value"], d["roofline"]["ms"] * 1e3))
        if d["value"] > best_v:
            best, best_v = envs[i], d["value"]
    except Exception as e:
        sys.stderr.write("variant %d failed: %s\n" % (i, e))
sys.stderr.write("default %.1f -> best: %s (%.1f)\n" % (base, best or "default", best_v))
print("export " + best if best else "true")
PY
cat $O/best_env_r2.sh
source $O/best_env_r2.sh

el "4. full GPU test-suite + bench under the selected configuration"
RECNN_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q --maxfail=12 --tb=short > $O/t_r2_best.log 2>&1
tail -4 $O/t_r2_best.log
timeout 150 python bench.py --steps 200 --warmup 20 > $O/bench_r2_best.json 2> $O/bench_r2_best.err
tail -c 600 $O/bench_r2_best.json
el "done"
