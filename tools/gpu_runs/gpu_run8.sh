set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "12 4" "12 5" "11 4"; do
set -- $cfg
echo "tile_bits=$1 low_bits=$2"
QSV_TRACE_PASSES=1 timeout 300 python bench.py --tile-bits $1 --low-bits $2 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>&1 >/dev/null | tail -6
done
