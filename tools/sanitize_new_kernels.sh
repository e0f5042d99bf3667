set -u
OUT=gpurun_out/r4b; mkdir -p $OUT
SEL="(test_row_stacked_kernels_vs_oracle and streams-auto) or uint8_target"
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py -m gpu -x -q -k "$SEL" > $OUT/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> $OUT/sanitizer_racecheck.log; tail -4 $OUT/sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py -m gpu -x -q -k "$SEL" > $OUT/sanitizer_synccheck.log 2>&1
echo "synccheck exit $?" >> $OUT/sanitizer_synccheck.log; tail -4 $OUT/sanitizer_synccheck.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py -m gpu -x -q -k "$SEL" > $OUT/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> $OUT/sanitizer_memcheck.log; tail -4 $OUT/sanitizer_memcheck.log
