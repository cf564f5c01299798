set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2_gputests_2gpu.log
cat gpurun_out/r2_gputests_2gpu.log
QSV_TRACE_PASSES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -8 gpurun_out/r2_bench_2gpu.err
tail -c 3000 gpurun_out/r2_bench_2gpu.json
