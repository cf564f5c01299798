timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
for sb in 268435456 1073741824 2147483648; do
echo "== staging $sb"
QSV_STAGING_BYTES=$sb timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['exchange'], d['max_abs_err_vs_closed_form'])"
done
