set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3h
( timeout 600 python -m pytest tests/test_gpu_sharded.py -q ) > ${O}_pytest2.log 2>&1
tail -3 ${O}_pytest2.log
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_variants.py -q -x ) > ${O}_pytest1.log 2>&1
tail -3 ${O}_pytest1.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for cfg in "QSV_PERM_ROUNDS=1" "QSV_PERM_ROUNDS=0"; do
  echo "== $cfg" >> ${O}_grover2.log; echo "== $cfg" >> ${O}_grover2.err
  env $cfg timeout 600 $TR bench.py --gpus 2 --workload grover --steps 2 --warmup 1 --no-cpu-baseline --no-extras >> ${O}_grover2.log 2>> ${O}_grover2.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2s3h_grover2.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    e=d['exchange']; print(d['ms_per_step'], [round(p['ms'],1) for p in d['roofline']['per_pass']], e['ms_per_step'], e['exposed_ms'], e['pipelined_remaps_per_step'], d.get('max_abs_err_vs_closed_form'))
PY
tail -3 ${O}_grover2.err
