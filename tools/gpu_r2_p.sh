# r02-p (N GPUs, N = $1): scaling evidence: torchrun bench on C1 and C3, in-library device list on C1
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for cfg in c1 c3; do
  steps=8; spp=64; [ $cfg = c3 ] && { steps=4; spp=8; }
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $steps --warmup 3 --config $cfg --spp-per-step $spp --no-cpu-baseline > gpurun_out/r02p_bench_${cfg}_${N}gpu.json 2> gpurun_out/bench_${cfg}.err
  tail -2 gpurun_out/bench_${cfg}.err | cut -c1-300
  python -c "
import json; d=json.loads(open('gpurun_out/r02p_bench_${cfg}_${N}gpu.json').read().strip().splitlines()[-1]); print('$cfg N=$N value %.1f e2e %.1f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step'])); print(json.dumps(d['ranks'])); print(json.dumps(d['gather_check']))"
done
timeout 600 python tools/bench_group.py --gpus $N > gpurun_out/r02p_group_${N}gpu.json 2> gpurun_out/group.err; cat gpurun_out/r02p_group_${N}gpu.json | cut -c1-1500; tail -2 gpurun_out/group.err
timeout 600 python -m pytest tests/test_gpu_bench_scenes.py tests/test_gpu_dropin.py -m gpu -q -k "multi_device or two_gpus" 2>&1 | tail -3
