# Final check of a round: full GPU test-suite, smoke, both bench arms on C1, one line per extra config.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 --ref-spp 16 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
for c in c2 c3 c4; do python bench.py --config $c --steps 4 --warmup 3 --spp-per-step 8 --no-cpu-baseline > gpurun_out/bench_$c.json 2>> gpurun_out/bench.err; done
python bench.py --impl reference --config c2 --steps 1 --warmup 0 --ref-spp 8 > gpurun_out/bench_c2_ref.json 2>> gpurun_out/bench.err
python - <<'PY'
import json
for f in ['bench_ref','bench','bench_c2','bench_c2_ref','bench_c3','bench_c4']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print('%-12s value %.2f e2e %.2f mrays %s launches %s' % (f, d['value'], d['e2e']['value'], d.get('mrays_per_s'), d.get('gpu_launches')))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench.err
