set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3m
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 8 --workload grover --steps 2 --warmup 1 --no-cpu-baseline --no-extras > ${O}_grover36.json 2> ${O}_grover36.err
python -c "
import json; d=json.load(open('${O}_grover36.json')); e=d['exchange']; print('grover36', d['ms_per_step'], [round(p['ms'],1) for p in d['roofline']['per_pass']], e['remaps_per_step'], e['ms_per_step'], e['exposed_ms'], e['pipelined_remaps_per_step'], d.get('max_abs_err_vs_closed_form'), d['norm_sqr'])"
tail -2 ${O}_grover36.err
timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline > ${O}_qft36.json 2> ${O}_qft36.err
python -c "
import json; d=json.load(open('${O}_qft36.json')); print('qft36', d['ms_per_step'], d['roofline']['frac'], d['sharded_parity_max_abs_err'], d['sharded_parity_remaps'], d['exchange_probe'], d.get('max_abs_err_vs_closed_form'))"
tail -2 ${O}_qft36.err
