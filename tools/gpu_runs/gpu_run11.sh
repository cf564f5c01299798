set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_gputests_2gpu.log
cat gpurun_out/r2_gputests_2gpu.log
for w in qft grover; do
QSV_TRACE_PASSES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --workload $w > gpurun_out/r2_bench_2gpu_$w.json 2> gpurun_out/r2_bench_2gpu_$w.err
grep -v "^\[qsv\]" gpurun_out/r2_bench_2gpu_$w.err | grep -i "error\|Traceback" | head -5
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_2gpu_$w.json').read().strip().splitlines()[-1])
    print('$w', d['ms_per_step'], [round(p['ms'],1) for p in d['roofline']['per_pass']][:12], d['exchange'], d.get('exchange_probe'), d['sharded_parity_max_abs_err'], d.get('max_abs_err_vs_closed_form'), d['e2e']['value'], d['config']['workload'][:80])
except Exception as e: print('no json', e)
PY
done
