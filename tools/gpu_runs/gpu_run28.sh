set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=$PWD/gpurun_out/r2s3l
for which in _ab_old . _ab_old . ; do
  cd $GRAFT_REPO_ROOT/$which
  echo "== $which" >> ${O}_ab.log
  QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>&1 | grep -E "^\[qsv\] pass|\"value\"" | tail -5 | cut -c1-130 >> ${O}_ab.log
done
cat ${O}_ab.log
