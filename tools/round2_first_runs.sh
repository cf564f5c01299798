#!/bin/bash
# First GPU runs of the next round (what round 1 ran out of GPU minutes for).  Each block is one gpurun call.
#   1 GPU : bash tools/round2_first_runs.sh one
#   2 GPUs: bash tools/round2_first_runs.sh two      (gpurun --gpus 2)
#   8 GPUs: bash tools/round2_first_runs.sh eight    (gpurun --gpus 8)
set -u
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 1), "ms  frac", round(d["roofline"]["frac"], 3), "err", d["max_abs_err_vs_closed_form"], d.get("exchange"))
PY
}
case "${1:-one}" in
one)
  # fused basis initialisation (pass_kernel_init.cu): parity, then timing of both modes against the default
  QSV_TEST_FUSED_INIT=1 timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q -k fused 2>&1 | tail -3
  for m in 0 1 2; do
    QSV_FUSED_INIT=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_fused$m.json; line gpurun_out/r2_fused$m.json
  done
  ;;
two)
  # sharded registers with the pipelined kernel on every pass (one-round passes included)
  for r in 2 1; do
    QSV_ASYNC_MIN_ROUNDS=$r timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_2gpu_minrounds$r.json; line gpurun_out/r2_2gpu_minrounds$r.json
  done
  ;;
eight)
  # peer-memory transport above 2 GPUs (never measured in round 1) against the NCCL transport
  for t in peer nccl; do
    if [ $t = nccl ]; then export QSV_NCCL_EXCHANGE=1; fi
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2_8gpu_$t.json; line gpurun_out/r2_8gpu_$t.json
  done
  ;;
esac
