#!/usr/bin/env bash
# Collects the evidence kept under profiles/ (run on a B200 box: gpurun -- 'bash tools/collect_profiles.sh TAG').
# Order matters: plain runs first (numbers), profiler runs afterwards (a number printed under ncu is never a bench value).
set -u
TAG=${1:-r01_final}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -1 $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
python tools/profile_pair.py --reps 5 > $OUT/${TAG}_pair_stage_times.log 2>&1; tail -1 $OUT/${TAG}_pair_stage_times.log | cut -c1-60
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; cut -c1-200 $OUT/${TAG}_bench_ref.json
python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
# launch list of one pair (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_pair.py --reps 3 > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_summary.txt 2>&1; head -12 $OUT/${TAG}_launch_summary.txt
# full captures of the top kernels (one launch each)
for K in match_tc_kernel replay_cells_kernel warp_fast_kernel harris_response_kernel replay_walk_bits_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o $OUT/${TAG}_$K \
      python tools/profile_pair.py --reps 1 > $OUT/ncu_$K.log 2>&1
done
ls -la $OUT/${TAG}_*
