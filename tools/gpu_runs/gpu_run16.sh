set -u
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r2_gputests_g.log
cat gpurun_out/r2_gputests_g.log
timeout 600 python tools/run_configs.py 30 2>&1 | grep "config 3" | head -2
