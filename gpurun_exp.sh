timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default2.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_default2.json').read()); print(d['ms_per_step'], d['roofline'], d['e2e'], d['config'])"
