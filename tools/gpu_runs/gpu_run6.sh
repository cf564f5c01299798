set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
QSV_FUSED_INIT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 5 -c 1 -o gpurun_out/r2_tma_lean_n32 \
   python bench.py --qubits 32 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log
