timeout 200 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/m_final2.log 2>&1
echo "rc=$?"
tail -1 gpurun_out/m_final2.log > gpurun_out/bench_2gpu_final.json
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_2gpu_final.json').read()); print('  ', d['ms_per_step'], d['exchange'], d['max_abs_err_vs_closed_form'], d['roofline']['frac'])" 2>/dev/null || (grep -v "^\s*File\|^\s*\^\|^    " gpurun_out/m_final2.log | tail -12 | cut -c1-300)
