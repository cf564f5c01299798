set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lb in 3 4 5; do
echo "low_bits=$lb"
QSV_FUSED_INIT=0 QSV_TRACE_PASSES=1 timeout 300 python bench.py --qubits 32 --low-bits $lb --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>&1 >/dev/null | tail -6
done
