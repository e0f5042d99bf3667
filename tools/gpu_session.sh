#!/bin/bash
# One gpurun call = tests + smoke + bench lines + ncu launch list + one ncu --set full capture.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_session.sh <tag> [parts]
#   parts: any of  tests smoke bench bench_more ref cudnn launches full   (default: all)
set -u
TAG=${1:-r1}
PARTS=${2:-"tests smoke bench bench_more ref cudnn peak launches full"}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ " $PARTS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > "$OUT/gpu.txt" 2>&1

if has tests; then
  timeout 1800 python -m pytest tests -m gpu -q -s ${PYTEST_ARGS:-} > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"; grep -E "math=|FAILED|passed|failed|Error" "$OUT/pytest_gpu.log" | tail -40
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
  echo "smoke exit $?" >> "$OUT/smoke.log"; tail -3 "$OUT/smoke.log"
fi
if has bench; then
  timeout 600 python bench.py > "$OUT/bench_espcn.json" 2> "$OUT/bench_espcn.err"
  echo "bench exit $?"; cat "$OUT/bench_espcn.json"
fi
if has bench_more; then
  for spec in "vdsr_b64_128 auto" "edsr64_x4_b32_lr32 auto" "srcnn_x2_b16 auto" "fsrcnn_x4_b16_lr32 auto" "srgan_x4_b16 auto" \
              "edsr256_x4_b32_lr32 bf16" "edsr256_x4_b32_lr32 auto"; do
    set -- $spec; wl=$1; mth=$2
    timeout 600 python bench.py --workload $wl --math $mth --steps 20 --warmup 5 --no-cpu-baseline --no-sub > "$OUT/bench_${wl}_$mth.json" 2> "$OUT/bench_${wl}_$mth.err"
    echo "bench $wl $mth exit $?"; cut -c1-420 "$OUT/bench_${wl}_$mth.json"; tail -2 "$OUT/bench_${wl}_$mth.err"
  done
fi
if has ref; then
  timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
  cat "$OUT/bench_reference.json"
fi
if has cudnn; then
  for wl in ${CUDNN_WL:-espcn_x4_b128_lr64 vdsr_b64_128 edsr64_x4_b32_lr32 edsr256_x4_b32_lr32}; do
    timeout 600 python bench.py --impl cudnn --workload $wl --steps 30 --warmup 5 > "$OUT/bench_cudnn_$wl.json" 2> "$OUT/bench_cudnn_$wl.err"
    echo "cudnn $wl exit $?"; cut -c1-400 "$OUT/bench_cudnn_$wl.json"
  done
fi
if has peak; then
  TF32_PEAK_OUT="$OUT/tf32_peak.json" timeout 120 python tools/measure_tf32_peak.py 2> "$OUT/tf32_peak.err"
fi
if has sanitize; then
  # mbarrier / TMEM hand-shake protocols under racecheck + synccheck: a slice of the op sweep
  SEL=${SAN_K:-"(test_fused_conv_vs_oracle and auto and True) or test_exact_mode_op or test_golden_net"}
  timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" > "$OUT/sanitizer_racecheck.log" 2>&1
  echo "racecheck exit $?" >> "$OUT/sanitizer_racecheck.log"; tail -4 "$OUT/sanitizer_racecheck.log"
  timeout 420 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" > "$OUT/sanitizer_synccheck.log" 2>&1
  echo "synccheck exit $?" >> "$OUT/sanitizer_synccheck.log"; tail -4 "$OUT/sanitizer_synccheck.log"
  timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" > "$OUT/sanitizer_memcheck.log" 2>&1
  echo "memcheck exit $?" >> "$OUT/sanitizer_memcheck.log"; tail -4 "$OUT/sanitizer_memcheck.log"
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file "$OUT/launches_espcn.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > "$OUT/launches_run.log" 2>&1
  echo "ncu launches exit $?"
fi
if has full; then
  # FULL_WL / FULL_K / FULL_SKIP / FULL_COUNT select the workload, kernel regex and launch window of the --set full capture
  WL=${FULL_WL:-espcn_x4_b128_lr64}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${FULL_K:-k_conv_rs|k_conv_sl|k_tc_wgrad}" \
      --launch-skip ${FULL_SKIP:-16} -c ${FULL_COUNT:-10} -f -o "$OUT/full_$WL" python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline --no-graph > "$OUT/full_run.log" 2>&1
  echo "ncu full exit $?"
  ls -la "$OUT"
fi
