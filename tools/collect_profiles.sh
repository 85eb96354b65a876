#!/usr/bin/env bash
# Collects the evidence kept under profiles/ (run on a B200 box: gpurun -- 'bash tools/collect_profiles.sh TAG').
# Order matters: plain runs first (numbers), profiler runs afterwards (a number printed under ncu is never a bench value).
set -u
TAG=${1:-r02_final}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -1 $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; cut -c1-240 $OUT/${TAG}_bench_ref.json
python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
python bench.py --workload c3 > $OUT/${TAG%_final}_c3_pair.json 2> $OUT/c3.err; cut -c1-400 $OUT/${TAG%_final}_c3_pair.json
python bench.py --workload c1 > $OUT/${TAG%_final}_c1_mountain.json 2> $OUT/c1.err; cut -c1-400 $OUT/${TAG%_final}_c1_mountain.json
python bench.py --workload c2 > $OUT/${TAG%_final}_c2_oilseed.json 2> $OUT/c2.err; cut -c1-400 $OUT/${TAG%_final}_c2_oilseed.json
python tools/cpu_o0.py > $OUT/${TAG%_final}_cpu_o0.json 2> $OUT/cpu_o0.err; cut -c1-300 $OUT/${TAG%_final}_cpu_o0.json
# launch list of one pair (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_pair.py --reps 3 > $OUT/ncu_launches.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_summary.txt 2>&1; head -14 $OUT/${TAG}_launch_summary.txt
# full captures of the kernels that changed (one launch each)
for K in match_tc_kernel warp_quad_kernel harris_fused_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o $OUT/${TAG}_$K \
      python tools/profile_pair.py --reps 1 > $OUT/ncu_$K.log 2>&1
done
python tools/ncu_metrics.py ${TAG} $OUT/${TAG}_match_tc_kernel.ncu-rep $OUT/${TAG}_warp_quad_kernel.ncu-rep $OUT/${TAG}_harris_fused_kernel.ncu-rep > /dev/null 2>&1
cp profiles/${TAG}_ncu_metrics.json $OUT/ 2>/dev/null
ls -la $OUT/${TAG}_* | head -30
