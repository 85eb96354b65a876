set -u
mkdir -p gpurun_out
bash tools/probe_box.sh > gpurun_out/r02_box.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_a_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_a_pytest_gpu.log
timeout 120 tools/microbench gpurun_out/r02_microbench.json > gpurun_out/r02_microbench.log 2>&1; cat gpurun_out/r02_microbench.json | cut -c1-1500
timeout 200 python tools/copy_ceiling.py --gpus 1 --numa 1 > gpurun_out/r02_copy_ceiling_1gpu.json 2> gpurun_out/r02_copy_ceiling.err; cat gpurun_out/r02_copy_ceiling_1gpu.json | cut -c1-600
timeout 200 python tools/copy_ceiling.py --gpus 1 --numa 0 >> gpurun_out/r02_copy_ceiling_1gpu.json 2>> gpurun_out/r02_copy_ceiling.err; tail -1 gpurun_out/r02_copy_ceiling_1gpu.json | cut -c1-400
timeout 300 python tools/profile_pair.py --reps 5 > gpurun_out/r02_a_pair_stage_times.log 2>&1; tail -1 gpurun_out/r02_a_pair_stage_times.log | cut -c1-400
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err; cut -c1-700 gpurun_out/r02_a_bench.json
head -60 gpurun_out/r02_box.txt
