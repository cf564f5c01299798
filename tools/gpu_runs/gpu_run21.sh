set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3d
nvidia-smi -L | head -3
( timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q ) > ${O}_pytest2.log 2>&1
tail -5 ${O}_pytest2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for cfg in "QSV_OVERLAP=1" "QSV_OVERLAP=0" "QSV_XCHG_SMS=16" "QSV_XCHG_SMS=36"; do
  echo "== $cfg" >> ${O}_grover2.log
  env $cfg timeout 600 $TR bench.py --gpus 2 --workload grover --steps 2 --warmup 1 --no-cpu-baseline --no-extras >> ${O}_grover2.log 2>> ${O}_grover2.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2s3d_grover2.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: continue
    print(d['ms_per_step'], d['exchange'], d.get('max_abs_err_vs_closed_form'))
PY
tail -5 ${O}_grover2.err
