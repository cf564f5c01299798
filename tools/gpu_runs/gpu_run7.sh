set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for q in 1 0; do
QSV_QFT4=$q QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_e$q.json 2> gpurun_out/r2_bench_e$q.err
echo "qft4=$q"; tail -4 gpurun_out/r2_bench_e$q.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_e$q.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['roofline']['frac'],d['max_abs_err_vs_closed_form'])"
done
(timeout 1200 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2_gputests_e.log
cat gpurun_out/r2_gputests_e.log
QSV_FUSED_INIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 5 -c 1 -o gpurun_out/r2_tma_qft4_n32 \
   python bench.py --qubits 32 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu3.log 2>&1
tail -2 gpurun_out/r2_ncu3.log
