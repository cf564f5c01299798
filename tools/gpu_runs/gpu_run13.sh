set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/run_configs.py 30 2>&1 | grep "config 3" > gpurun_out/r2_configs3.txt
cat gpurun_out/r2_configs3.txt
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_variants.py -m gpu -x -q 2>&1 | tail -4)
QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>&1 >/dev/null | tail -4
