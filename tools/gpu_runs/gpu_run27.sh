set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=$PWD/gpurun_out/r2s3j
for rep in 1 2; do
  for which in new old; do
    if [ $which = old ]; then cd $GRAFT_REPO_ROOT/_ab_old; else cd $GRAFT_REPO_ROOT; fi
    echo "== $which rep $rep" >> ${O}_ab.log
    QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>&1 | grep -E "^\[qsv\] pass|\"value\"" | tail -5 | cut -c1-160 >> ${O}_ab.log
  done
done
cat ${O}_ab.log
