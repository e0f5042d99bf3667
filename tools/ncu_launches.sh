#!/bin/bash
# ncu launch list (gpu__time_duration per kernel launch) of one bench.py configuration, summarised per kernel name.
# Usage: bash tools/ncu_launches.sh <tag> <bench args...>
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-1500} --csv \
    --log-file "$OUT/launches.csv" python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-sub > "$OUT/launches_run.log" 2>&1
echo "ncu exit $?"
python tools/launch_summary.py "$OUT/launches.csv" ${NCU_SKIP:-0} | tee "$OUT/launch_summary.txt"
