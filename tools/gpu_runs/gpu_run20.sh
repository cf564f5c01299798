set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3c
run() { echo "== $*" >> ${O}_config3.log; env "$@" QSV_TRACE_PASSES=1 timeout 300 python tools/config3_probe.py 30 100 >> ${O}_config3.log 2>&1; }
run A=1
run QSV_WARP_LOCAL=0
run QSV_LIB=$PWD/quantr_b200/libqsv_p0.so
run QSV_LOW_BITS=5
run QSV_LOW_BITS=5 QSV_WARP_LOCAL=0
run QSV_LOW_BITS=6
run QSV_LOW_BITS=6 QSV_WARP_LOCAL=0
run QSV_LOW_BITS=5 QSV_LIB=$PWD/quantr_b200/libqsv_p0.so
grep -E "^==|^rep 1" ${O}_config3.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
tail -3 ${O}_pytest.log
