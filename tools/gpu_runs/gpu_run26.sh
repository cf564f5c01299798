set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3i
( timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
tail -3 ${O}_pytest.log
timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
python -c "
import json; d=json.load(open('${O}_bench_default.json')); print(d['value'], d['roofline']['frac'], d['roofline']['min_pass_frac'], [round(p['ms'],1) for p in d['roofline']['per_pass']], d['config3']['ms'], d['e2e']['value'], d['small_configs'])"
tail -2 ${O}_bench_default.err
timeout 300 python bench.py --workload grover --qubits 31 --steps 2 --warmup 1 --no-cpu-baseline --no-extras > ${O}_grover31.json 2> ${O}_grover31.err
python -c "
import json; d=json.load(open('${O}_grover31.json')); print(d['value'], [round(p['ms'],1) for p in d['roofline']['per_pass']], d.get('max_abs_err_vs_closed_form'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_qft33.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > ${O}_ncu_bench.log 2>&1
tail -2 ${O}_launches_qft33.csv | cut -c1-200
