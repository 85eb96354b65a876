set -u
mkdir -p gpurun_out
runN() { N=$1; tag=$2; shift; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --pairs 128 --no-cpu > gpurun_out/r02_j_bench_$tag.json 2> gpurun_out/r02_j_bench_$tag.err; python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02_j_bench_$tag.json') if l.startswith('{')][-1]
    print('$tag', 'N=%d value %.0f ms/pair/gpu %.3f | e2e %.0f (frac %.2f of ceiling %.0f) | lanes %s cores/rank %s launches/pair %.0f' % (d['n_gpus'], d['value'], d['ms_per_pair'], d['e2e']['value'], d['e2e']['copy_ceiling']['frac'], d['e2e']['copy_ceiling']['value'], d['config']['lanes'], d['config']['host_threads_this_rank'], d['gpu_launches_per_pair']))
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02_j_bench_$tag.err').read()[-2000:])
PY
}
runN 8 n8
runN 8 n8_l5 PANO_BATCH_LANES=5
runN 4 n4
for N in 8 4 2 1; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --workload chain --gpus $N --steps 4 --warmup 2 > gpurun_out/r02_chain_${N}gpu.json 2> gpurun_out/r02_chain_${N}gpu.err; python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02_chain_${N}gpu.json') if l.startswith('{')][-1]
    print('chain N=%d %.2f ms %.0f MP/s identical %s sha %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['identical_to_single_gpu'], d['canvas_sha'][:12]))
except Exception as e:
    print('chain $N FAILED', e); print(open('gpurun_out/r02_chain_${N}gpu.err').read()[-1500:])
PY
done
