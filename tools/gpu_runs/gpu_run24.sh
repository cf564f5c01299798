set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3g
( timeout 900 python -m pytest tests/test_gpu_sharded.py -q ) > ${O}_pytest2.log 2>&1
tail -30 ${O}_pytest2.log
