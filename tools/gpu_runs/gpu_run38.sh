set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3v
timeout 140 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err
python -c "
import json; d=json.load(open('${O}_bench.json')); print(d['value'], d['roofline']['frac'], d['without_prefix_folding'], d['config3']['ms'], d['small_configs']['config2_qft16_simulate_get_state_us'], d['e2e']['value'])"
tail -3 ${O}_bench.err
