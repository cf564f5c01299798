timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { lab=$1; shift; envs=$1; shift; echo "== $lab"; env QSV_TRACE_PASSES=1 $envs timeout 300 python bench.py --no-cpu-baseline "$@" 2>&1 | grep -E "^\[qsv\]|ms_per_step" | tail -6 | sed -E 's/.*"ms_per_step": ([0-9.]+).*max_abs_err_vs_closed_form": ([0-9.e-]+).*/ms_per_step \1 err \2/' | cut -c1-120; }
run q33_t11_l3 "A=1" --steps 2 --warmup 1 --tile-bits 11 --low-bits 3
run q33_t11_auto "A=1" --steps 2 --warmup 1 --tile-bits 11
run q33_t12_l4 "A=1" --steps 2 --warmup 1 --tile-bits 12 --low-bits 4
run q33_t12_auto "A=1" --steps 2 --warmup 1 --tile-bits 12
