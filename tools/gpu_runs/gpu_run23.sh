set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3f
( timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
tail -3 ${O}_pytest.log
run() { echo "== $*" >> ${O}_config3.log; env "$@" timeout 300 python tools/config3_probe.py 30 100 >> ${O}_config3.log 2>&1; }
run QSV_LOW_BITS=5
run QSV_LOW_BITS=5 QSV_WARP_LOCAL=0
run QSV_LOW_BITS=6
run QSV_LOW_BITS=6 QSV_WARP_LOCAL=0
run A=1
grep -E "^==|^rep 1" ${O}_config3.log
timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
python -c "
import json; d=json.load(open('${O}_bench_default.json')); print(d['value'], d['roofline']['frac'], d['roofline']['min_pass_frac'], [round(p['ms'],1) for p in d['roofline']['per_pass']], d['config3'], d['e2e']['value'])"
tail -2 ${O}_bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 5 -c 2 -o ${O}_qft32 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --qubits 32 > ${O}_ncu_qft32.log 2>&1
tail -2 ${O}_ncu_qft32.log
ls -la gpurun_out | tail -5
