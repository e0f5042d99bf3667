# compute-sanitizer over the kernels added late in round 2: two-stream k_conv_rs, rows+co-stacked wgrad (flag 524288 in the
# one-stream variant of the sweep), uint8-target loss epilogue.  Usage (GPU box): bash tools/sanitize_new_kernels.sh [outdir]
set -u
OUT=${1:-gpurun_out/sanitize}; mkdir -p $OUT
SEL="test_row_stacked_kernels_vs_oracle or uint8_target"
for tool in racecheck synccheck memcheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py -m gpu -x -q -k "$SEL" > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> $OUT/sanitizer_$tool.log; tail -4 $OUT/sanitizer_$tool.log
done
