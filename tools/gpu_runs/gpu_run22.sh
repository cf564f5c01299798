set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3e
( timeout 600 python -m pytest tests/test_gpu_sharded.py -q ) > ${O}_pytest2.log 2>&1
tail -5 ${O}_pytest2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for cfg in "QSV_OVERLAP=1" "QSV_OVERLAP=0"; do
  echo "== $cfg" >> ${O}_grover2.log; echo "== $cfg" >> ${O}_grover2.err
  env $cfg QSV_TRACE_EXCHANGE=1 timeout 600 $TR bench.py --gpus 2 --workload grover --steps 1 --warmup 1 --no-cpu-baseline --no-extras >> ${O}_grover2.log 2>> ${O}_grover2.err
done
grep -E "^==|remap:" ${O}_grover2.err
