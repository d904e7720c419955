# r02-zb (1 GPU): measured parity figures of every GPU-vs-oracle test (to set the bars from data)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --tb=short 2>&1 | grep -E "frac within|exact|passed|failed|::" > gpurun_out/r02zb_parity_figures.log; tail -3 gpurun_out/r02zb_parity_figures.log
