# r02-a: round-1 kernels + the new bench-scene / shard parity tests + the fixed reference arm (baseline for round 2)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
nproc; python -c "import bench, json; print(json.dumps(bench.host_cpu_info()))"
python -m pytest tests/test_gpu_bench_scenes.py -m gpu -x -q -s 2>&1 | tail -40
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_bench_scenes.py 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2>> gpurun_out/bench.err; cut -c1-1500 gpurun_out/r02a_bench_ref.json
python bench.py --steps 8 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/bench.err; cat gpurun_out/r02a_bench.json; tail -5 gpurun_out/bench.err
