set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for il in 1 0; do
QSV_TILE_INTERLEAVE=$il QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_d$il.json 2> gpurun_out/r2_bench_d$il.err
echo "interleave=$il"; tail -4 gpurun_out/r2_bench_d$il.err
done
(timeout 1200 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2_gputests_d.log
QSV_FUSED_INIT=0 timeout 600 python tools/stream_probe.py 32 > gpurun_out/r2_stream_probe_interleave.txt 2>&1
head -8 gpurun_out/r2_stream_probe_interleave.txt
