set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3r
( timeout 600 python -m pytest tests/test_gpu_sharded.py -q ) > ${O}_pytest2.log 2>&1
tail -3 ${O}_pytest2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > ${O}_qft34.json 2> ${O}_qft34.err
timeout 600 $TR bench.py --gpus 2 --workload grover --steps 2 --warmup 1 --no-cpu-baseline --no-extras > ${O}_grover34.json 2> ${O}_grover34.err
python - <<'PY'
import json
for f in ('qft34','grover34'):
    lines=[l for l in open(f'gpurun_out/r2s3r_{f}.json') if l.startswith('{')]
    d=json.loads(lines[-1]); e=d['exchange']
    print(f, d['ms_per_step'], round(d['roofline']['frac'],3), [round(p['ms'],1) for p in d['roofline']['per_pass']], e['remaps_per_step'], e['exposed_ms'], e['pipelined_remaps_per_step'], d.get('max_abs_err_vs_closed_form'), d.get('sharded_parity_max_abs_err'), d['prefix_ops_folded_into_initial_state'])
PY
tail -2 ${O}_grover34.err
