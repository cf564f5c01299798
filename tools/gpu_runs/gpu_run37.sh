set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 4 --steps 2 --warmup 1 --no-cpu-baseline > ${O}_qft35.json 2> ${O}_qft35.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r2s3u_qft35.json') if l.startswith('{')]
d=json.loads(lines[-1]); e=d['exchange']
print('qft35', d['ms_per_step'], round(d['roofline']['frac'],3), [round(p['ms'],1) for p in d['roofline']['per_pass']], d.get('max_abs_err_vs_closed_form'), d.get('sharded_parity_max_abs_err'), d['prefix'], d['exchange_probe']['ms'] if d.get('exchange_probe') else None)
PY
tail -2 ${O}_qft35.err
