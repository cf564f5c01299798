timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # name, env..., args
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BARGS 2>&1 | tail -1 > gpurun_out/exp_$name.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/exp_$name.json').read()); print('$name', round(d['ms_per_step'],1), 'avg pass', round(d['roofline']['avg_launch_ms'],1), 'frac', round(d['roofline']['frac'],3), 'err', d['max_abs_err_vs_closed_form'])" || tail -3 gpurun_out/exp_$name.json
}
BARGS="" run t12_fast A=1
BARGS="" run t12_nofast QSV_NO_FAST=1
BARGS="--tile-bits 11" run t11_fast A=1
BARGS="--tile-bits 11" run t11_nofast QSV_NO_FAST=1
BARGS="--tile-bits 11" run t11_occ5 QSV_LIB=/root/repo/quantr_b200/libqsv_occ5.so
BARGS="--tile-bits 11" run t11_occ5_nofast QSV_LIB=/root/repo/quantr_b200/libqsv_occ5.so QSV_NO_FAST=1
