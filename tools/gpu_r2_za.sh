# r02-za (1 GPU): RoughCoatBsdf: GPU vs oracle, vs the reference's render, through the drop-in; C2 sanity (unchanged lobes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -q --tb=short -k "coat or golden or material or dropin_writes or reference_binary" 2>&1 | grep -v "^$" | tail -6
python bench.py --config c2 --steps 3 --warmup 3 --spp-per-step 8 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
