timeout 90 python tools/single_h_probe.py 33 24 2>&1 | tail -4
echo "probe rc=$?"
timeout 240 python -m pytest tests -m gpu -x -q -k "variants or closed_form or scale" 2>&1 | tail -2
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-250
