set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3b
( timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
tail -3 ${O}_pytest.log
for cfg in "default" "QSV_MERGE_CTRL=0" "LB=3" "LB=5"; do
  echo "== $cfg" >> ${O}_config3.log
  case "$cfg" in
    default) QSV_TRACE_PASSES=1 timeout 300 python tools/config3_probe.py 30 100 >> ${O}_config3.log 2>&1 ;;
    QSV_MERGE_CTRL=0) QSV_MERGE_CTRL=0 QSV_TRACE_PASSES=1 timeout 300 python tools/config3_probe.py 30 100 >> ${O}_config3.log 2>&1 ;;
    LB=3) QSV_LOW_BITS=3 QSV_TRACE_PASSES=1 timeout 300 python tools/config3_probe.py 30 100 >> ${O}_config3.log 2>&1 ;;
    LB=5) QSV_LOW_BITS=5 QSV_TRACE_PASSES=1 timeout 300 python tools/config3_probe.py 30 100 >> ${O}_config3.log 2>&1 ;;
  esac
done
grep -E "^==|^rep" ${O}_config3.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > ${O}_bench.json 2> ${O}_bench.err
python -c "
import json; d=json.load(open('${O}_bench.json')); print(d['value'], d['roofline']['frac'], [round(p['ms'],1) for p in d['roofline']['per_pass']])"
QSV_FUSED_INIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 8 -c 1 -o ${O}_config3_n28 python tools/config3_probe.py 28 10 > ${O}_ncu_c3.log 2>&1
tail -2 ${O}_ncu_c3.log
