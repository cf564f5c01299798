set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2_gputests_c.log
for st in 0 1; do
QSV_TMA_STORE=$st QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_c$st.json 2> gpurun_out/r2_bench_c$st.err
echo "tma_store=$st"; tail -4 gpurun_out/r2_bench_c$st.err
done
(QSV_TMA_STORE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "qft_closed_form or layered or config3_depth100_live" 2>&1 | tail -3) >> gpurun_out/r2_gputests_c.log
QSV_FUSED_INIT=0 timeout 600 python tools/stream_probe.py 32 > gpurun_out/r2_stream_probe.txt 2>&1
cat gpurun_out/r2_stream_probe.txt
