# r02-zd (1 GPU): confirm the two re-barred tests at HEAD
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -k "many_lights or many_analytic" 2>&1 | tail -3
