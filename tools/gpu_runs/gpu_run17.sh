set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
QSV_FUSED_INIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tma -s 8 -c 1 -o gpurun_out/r2_config3_n28 python tools/config3_probe.py 28 10 > gpurun_out/r2_ncu4.log 2>&1
tail -3 gpurun_out/r2_ncu4.log
