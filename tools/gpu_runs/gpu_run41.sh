set -u
cd $GRAFT_REPO_ROOT
timeout 40 python -m pytest tests/test_gpu_parity.py -q -k "wide_multi or x3sudoko" 2>&1 | tail -2
timeout 30 python -m pytest tests/test_cpp_host.py -m gpu -q 2>&1 | tail -2
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
