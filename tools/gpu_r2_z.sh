# r02-z (2 GPUs): final code: multi-device tests (library + drop-in) and the 2-rank bench with gather check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_scenes.py tests/test_gpu_dropin.py -m gpu -q -k "multi_device or two_gpus" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02z_bench_2gpu.json 2> gpurun_out/bench2.err
tail -1 gpurun_out/bench2.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/r02z_bench_2gpu.json').read().strip().splitlines()[-1]); print('N=2 value %.1f e2e %.1f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'])); print(json.dumps(d['ranks'])); print(json.dumps(d['gather_check']))"
