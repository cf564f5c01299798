set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2s3_pytest.log 2>&1
tail -4 gpurun_out/r2s3_pytest.log
timeout 600 python bench.py > gpurun_out/r2s3_bench_default.json 2> gpurun_out/r2s3_bench_default.err
tail -c 600 gpurun_out/r2s3_bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s3_launches_qft33.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2s3_ncu_bench.log 2>&1
tail -3 gpurun_out/r2s3_launches_qft33.csv
QSV_FUSED_INIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 8 -c 1 -o gpurun_out/r2s3_config3_n28 python tools/config3_probe.py 28 10 > gpurun_out/r2s3_ncu_c3.log 2>&1
tail -3 gpurun_out/r2s3_ncu_c3.log
QSV_TRACE_PASSES=1 timeout 300 python tools/config3_probe.py 30 100 > gpurun_out/r2s3_config3_n30.log 2>&1
tail -3 gpurun_out/r2s3_config3_n30.log
ls -la gpurun_out
