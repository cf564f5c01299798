set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2s3t
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 2 -c 2 -o ${O}_qft32 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --qubits 32 > ${O}_ncu_qft32.log 2>&1
tail -2 ${O}_ncu_qft32.log
ls -la ${O}_qft32.ncu-rep
