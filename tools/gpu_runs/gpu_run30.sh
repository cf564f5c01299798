set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3n
run() { echo "== $*" >> ${O}_config3.log; env "$@" timeout 300 python tools/config3_probe.py 30 100 >> ${O}_config3.log 2>&1; }
run A=1
run QSV_NO_LEAN=1
run A=1
run QSV_NO_LEAN=1
grep -E "^==|^rep 1" ${O}_config3.log
( timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
tail -3 ${O}_pytest.log
