set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3p
QSV_TRACE_PASSES=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>&1 | grep -E "^\[qsv\] pass|\"value\"|rror" | tail -6 | cut -c1-300
( timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1
tail -3 ${O}_pytest.log
timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
python -c "
import json; d=json.load(open('${O}_bench_default.json')); print(d['value'], d['roofline']['frac'], [(round(p['ms'],1), p['kind'][:5]) for p in d['roofline']['per_pass']], d['config3']['ms'], d['e2e']['value'], d['max_abs_err_vs_closed_form'], d['norm_sqr'], d['prefix_ops_folded_into_initial_state'])"
tail -2 ${O}_bench_default.err
