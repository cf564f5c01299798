timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default3.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_default3.json').read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1
tail -3 gpurun_out/launches_r01_final2.csv | cut -c1-200
