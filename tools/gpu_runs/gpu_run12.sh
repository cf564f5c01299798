set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/run_configs.py 30 > gpurun_out/r2_configs.txt 2>&1
cat gpurun_out/r2_configs.txt
(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4)
