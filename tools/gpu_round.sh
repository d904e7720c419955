# One GPU session: parity tests, smoke, bench (both arms), ncu launch list + full capture of the hot kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 --ref-spp 16 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json | cut -c1-300
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --config c3 --steps 4 --warmup 3 --spp-per-step 8 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err
python bench.py --config c4 --steps 4 --warmup 3 --spp-per-step 8 --no-cpu-baseline > gpurun_out/bench_c4.json 2>> gpurun_out/bench.err
python bench.py --impl reference --config c4 --steps 1 --warmup 0 --ref-spp 4 > gpurun_out/bench_c4_ref.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for k in k_trace k_shadow_bvh; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -12
