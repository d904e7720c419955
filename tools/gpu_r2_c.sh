# r02-c: ncu of the rewritten kernels: launch list + full capture of k_trace for three node variants
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_q4.csv $B > gpurun_out/ncu_bench.log 2>&1
TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_f32.csv $B > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/r02c_k_trace_q4_t320 $B > gpurun_out/ncu1.log 2>&1
TGB_TREELET=0 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/r02c_k_trace_q4_t0 $B > gpurun_out/ncu2.log 2>&1
TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so ncu --set full --clock-control none --import-source on -k regex:k_trace -s 6 -c 1 -f -o gpurun_out/r02c_k_trace_f32 $B > gpurun_out/ncu3.log 2>&1
for k in k_shade k_accum k_regen; do
TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/r02c_$k $B > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -12
