mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
run() { # label, env, args
  lab=$1; shift; envs=$1; shift
  env $envs timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$lab.log 2>&1
  tail -1 gpurun_out/bench_$lab.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lab', 'ms/step', round(d['ms_per_step'],2), 'passes', d['fused_passes'], 'ms/pass', round(d['roofline']['avg_launch_ms'],2), 'frac', round(d['roofline']['frac'],3), 'err', d['max_abs_err_vs_closed_form'])" || tail -3 gpurun_out/bench_$lab.log
}
run q30_t12 "A=1" --qubits 30 --steps 3 --warmup 1
run q30_t11 "A=1" --qubits 30 --steps 3 --warmup 1 --tile-bits 11
run q30_t13 "A=1" --qubits 30 --steps 3 --warmup 1 --tile-bits 13
run q33_t12 "A=1" --steps 3 --warmup 1
run q33_t11 "A=1" --steps 3 --warmup 1 --tile-bits 11
run q33_t13 "A=1" --steps 3 --warmup 1 --tile-bits 13
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 3 -c 3 -o gpurun_out/prof_r01c python bench.py --qubits 30 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-100
