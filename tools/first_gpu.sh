nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25
python tools/perf_probe.py 3 512 8 2>&1 | tail -8
python tools/perf_probe.py 7 1024 8 2>&1 | tail -8
