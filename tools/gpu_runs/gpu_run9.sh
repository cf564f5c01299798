set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
tail -4 gpurun_out/r2_bench_f.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['roofline']['frac'],d['roofline']['min_pass_frac'],d['max_abs_err_vs_closed_form'],d['e2e']['value'])"
(timeout 1200 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2_gputests_f.log
cat gpurun_out/r2_gputests_f.log
