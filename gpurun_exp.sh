QSV_ASYNC=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/run_configs.py 30 2>&1 | grep -v Warning | tail -8
