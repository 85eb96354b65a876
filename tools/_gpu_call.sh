set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_i_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_i_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 --pairs 128 --no-cpu > gpurun_out/r02_i_bench.json 2> gpurun_out/r02_i_bench.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_i_bench.json'))
    print('default', 'value %.0f ms/pair %.3f | e2e %.0f (%.3f ms, frac %.2f of ceiling %.0f) | lat %.3f | warp %.1f us frac %.3f match %.1f us harris %.1f us | lanes %s depth %s launches/pair %.0f' % (d['value'], d['ms_per_pair'], d['e2e']['value'], d['e2e']['ms_per_pair'], d['e2e']['copy_ceiling']['frac'], d['e2e']['copy_ceiling']['value'], d['latency']['ms_per_pair_device_median'], 1e3*d['roofline']['kernel_ms'], d['roofline']['frac'], 1e3*d['roofline']['other_kernels']['match_tc_kernel']['ms'], 1e3*d['roofline']['other_kernels']['harris_fused_kernel']['ms_per_image'], d['config']['lanes'], d['config']['pipeline_depth'], d['gpu_launches_per_pair']))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r02_i_bench.err').read()[-1500:])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_tc_kernel -c 1 -o gpurun_out/r02_match_tc_kernel_v3 python tools/profile_pair.py --reps 1 > gpurun_out/ncu_match3.log 2>&1; tail -1 gpurun_out/ncu_match3.log
ncu -i gpurun_out/r02_match_tc_kernel_v3.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]; d=dict(zip(h,v))
for k in ('gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed.avg.per_cycle_elapsed','smsp__inst_executed.sum'): print(k, d.get(k))
"
