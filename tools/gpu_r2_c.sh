# r02-c: ncu of the rewritten kernels: launch lists + full captures exported to CSV on the box (the .ncu-rep files are too big to bring back)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_q4.csv $B > gpurun_out/ncu_bench.log 2>&1
TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_f32.csv $B > gpurun_out/ncu_bench.log 2>&1
cap() {  # name kernel-regex [env...]
  name=$1; k=$2; shift 2
  env "$@" ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o /tmp/$name $B > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep
}
cap r02c_k_trace_q4_t320 k_trace A=1
cap r02c_k_trace_q4_t0 k_trace TGB_TREELET=0
cap r02c_k_trace_f32 k_trace TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so
cap r02c_k_shade k_shade TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so
cap r02c_k_accum k_accum TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so
cap r02c_k_regen k_regen TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so
cap r02c_k_shadow_bvh k_shadow_bvh TGB200_LIB=$PWD/tungsten_b200/libtgb200_f32nodes.so
du -sh gpurun_out
