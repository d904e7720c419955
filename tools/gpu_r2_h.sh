# r02-h (2 GPUs): in-library multi-GPU context vs one GPU, 2-rank bench with gather check
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_bench_scenes.py -m gpu -q -k "multi_device or shard" 2>&1 | tail -8
python tools/bench_group.py --gpus 2 > gpurun_out/r02h_group_2gpu.json 2> gpurun_out/group.err; cat gpurun_out/r02h_group_2gpu.json; tail -3 gpurun_out/group.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/bench2.err; tail -2 gpurun_out/bench2.err; python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_2gpu.json').read().strip().splitlines()[-1]); print('N=2 value %.1f e2e %.1f' % (d['value'], d['e2e']['value'])); print(json.dumps(d['ranks'])); print(json.dumps(d['gather_check']))"
