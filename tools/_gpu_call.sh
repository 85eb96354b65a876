set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_c_pytest_gpu.log
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --pairs 128 --no-cpu > gpurun_out/r02_c_bench_$tag.json 2> gpurun_out/r02_c_bench_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_c_bench_$tag.json'))
    print('$tag', 'value %.0f ms/pair %.3f | e2e %.0f (%.3f ms, frac %.2f of ceiling %.0f) | lat %.3f | warp %.1f us frac %.3f | lanes %s depth %s launches/pair %.0f gen %.0fs' % (d['value'], d['ms_per_pair'], d['e2e']['value'], d['e2e']['ms_per_pair'], d['e2e']['copy_ceiling']['frac'], d['e2e']['copy_ceiling']['value'], d['latency']['ms_per_pair_device_median'], 1e3*d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['lanes'], d['config']['pipeline_depth'], d['gpu_launches_per_pair'], d['config']['generation_s']))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02_c_bench_$tag.err').read()[-1500:])
PY
}
run v2_d2_l7 PANO_BATCH_DEPTH=2 PANO_BATCH_LANES=7
run v1_d0_l16 PANO_BATCH_DEPTH=0 PANO_BATCH_LANES=16
run v3_d3_l6 PANO_BATCH_DEPTH=3 PANO_BATCH_LANES=6
run v4_d2_l12 PANO_BATCH_DEPTH=2 PANO_BATCH_LANES=12
run v5_d2_l7_chunked PANO_BATCH_DEPTH=2 PANO_BATCH_LANES=7 PANO_BATCH_REPLAY=0
run v6_d4_l7 PANO_BATCH_DEPTH=4 PANO_BATCH_LANES=7
python -c "
import json; d=json.load(open('gpurun_out/r02_c_bench_v2_d2_l7.json')); print(json.dumps(d['roofline'])[:1800]); print(d['latency'])"
timeout 300 python bench.py --workload c3 --no-cpu > gpurun_out/r02_c_c3.json 2> gpurun_out/r02_c_c3.err; cut -c1-900 gpurun_out/r02_c_c3.json; tail -3 gpurun_out/r02_c_c3.err
