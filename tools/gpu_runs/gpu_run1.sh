set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_gputests_a.log
for m in 0 2; do
QSV_FUSED_INIT=$m QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_tma_fused$m.json 2> gpurun_out/r2_tma_fused$m.err
tail -c 600 gpurun_out/r2_tma_fused$m.err
done
