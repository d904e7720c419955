mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 2 -o gpurun_out/prof_trace3 python bench.py --steps 1 --warmup 1 --spp-per-step 2 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | tail -3
