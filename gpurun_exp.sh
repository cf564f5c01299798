timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
for mode in peer nccl; do
echo "== $mode"
if [ $mode = nccl ]; then export QSV_NCCL_EXCHANGE=1; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench2_$mode.json
python -c "import json,sys; d=json.loads(open('gpurun_out/bench2_$mode.json').read()); print(d['ms_per_step'], d['exchange'], d['max_abs_err_vs_closed_form'])" || tail -5 gpurun_out/bench2_$mode.json
done
