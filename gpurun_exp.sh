mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --qubits 31 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_2gpu_q31.log 2>&1; tail -2 gpurun_out/bench_2gpu_q31.log | cut -c1-1800
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_2gpu_q34.log 2>&1; tail -2 gpurun_out/bench_2gpu_q34.log | cut -c1-1800
