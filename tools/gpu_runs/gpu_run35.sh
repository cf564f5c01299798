set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3s
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_qft33.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > ${O}_ncu_bench.log 2>&1
tail -2 ${O}_launches_qft33.csv | cut -c1-200
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras > ${O}_bench.json 2> ${O}_bench.err
python -c "
import json; d=json.load(open('${O}_bench.json')); print(d['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['prefix'], d['gpu_launches'])"
