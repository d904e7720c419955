# r02-zc (1 GPU): the tightened parity bars + cube_city loop-vs-BVH
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --tb=short 2>&1 | grep -E "cube_city, per-ray|passed|failed|Error|assert" | tail -8
