# one --set full capture of each hot kernel on a full (4 Mi path) mid-step launch
mkdir -p gpurun_out
for k in k_shade k_accum k_shadow_bvh k_trace k_shadow_prep; do
  ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 6 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -8
