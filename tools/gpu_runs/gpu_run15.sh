set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in qft grover; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2 --warmup 2 --no-cpu-baseline --workload $w > gpurun_out/r2_bench_8gpu_$w.json 2> gpurun_out/r2_bench_8gpu_$w.err
grep -i "error\|Traceback" gpurun_out/r2_bench_8gpu_$w.err | head -5
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_8gpu_$w.json').read().strip().splitlines()[-1])
    print('$w', d['ms_per_step'], [round(p['ms'],1) for p in d['roofline']['per_pass']][:30], d['exchange'], d.get('exchange_probe'), d['sharded_parity_max_abs_err'], d.get('max_abs_err_vs_closed_form'), d['norm_sqr'], d['e2e']['value'])
except Exception as e: print('no json', e)
PY
done
