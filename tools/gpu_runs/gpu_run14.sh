set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q 2>&1 | tail -6)
