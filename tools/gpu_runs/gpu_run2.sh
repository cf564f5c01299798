set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sampling_large or chi_square or failed_collapse or golden" 2>&1 | tail -5) > gpurun_out/r2_gputests_b.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -c 1500 gpurun_out/r2_bench_b.err
# ncu: full capture of two 2-round passes at n=32 (no fused init so that every launch is an ordinary pass)
QSV_FUSED_INIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 5 -c 2 -o gpurun_out/r2_tma_n32 \
   python bench.py --qubits 32 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
