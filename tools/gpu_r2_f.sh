# r02-f: f3 adaptive sampling + resume through the drop-in, full parity suites, C1 bench with the streaming-kernel rooflines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q -k "resume or adaptive" 2>&1 | tail -60
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dropin.py::test_dropin_resume_render_continues_bit_exactly 2>&1 | tail -8
python bench.py --steps 8 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f e2e %.1f trace %.0f ms (%.0f Mq/s) shadow %.0f ms (%.0f Mq/s) dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['mqueries_per_s'], r['k_shadow']['kernel_ms'], r['k_shadow']['mqueries_per_s'], d['device_ms']))
print(json.dumps(d['roofline_streaming'])); print(json.dumps(d['loop'])); print(json.dumps(d['other_configs'])); print(json.dumps(d['setup'])); print(d['timing'])"
