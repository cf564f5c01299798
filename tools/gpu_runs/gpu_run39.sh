set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python bench.py > gpurun_out/r2s3w_bench_default.json 2> gpurun_out/r2s3w_bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2s3w_bench_default.json')); print(d['value'], d['roofline']['frac'], d['without_prefix_folding']['ms_per_step'], d['cpu_baseline']['value'], d['same_config']['speedup'])"
