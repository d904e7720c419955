# r02-r (1 GPU): k_shade occupancy A/B on C1 (all-Lambert instantiation), C2 breakdown + ncu of the material-sorted k_shade
mkdir -p gpurun_out
run() { python bench.py --config $2 --steps $3 --warmup 3 --spp-per-step $4 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['roofline_streaming']
print('$1 $2: value %.1f e2e %.1f trace %.0f ms shadow %.0f ms shade %.0f ms accum %.0f ms regen %.0f ms prep %.0f sort %.0f dev %.0f ms' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['k_shadow']['kernel_ms'], s['k_shade']['kernel_ms'], s['k_accum']['kernel_ms'], s['k_regen']['kernel_ms'], d['loop']['k_shadow_prep_ms'], d['loop']['sort_ms'], d['device_ms']))"; }
for v in base shade7 shade6 shade10; do
  if [ "$v" = base ]; then unset TGB200_LIB; else export TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so; fi
  run $v c1 4 64
done
unset TGB200_LIB
run base c2 3 8; run base c3 3 8
ncu --set full --clock-control none --import-source on -k 'regex:k_shade' -s 5 -c 1 -f -o /tmp/s python bench.py --steps 1 --warmup 1 --spp-per-step 8 --no-cpu-baseline --no-other-configs --config c2 > gpurun_out/ncu_s.log 2>&1
ncu -i /tmp/s.ncu-rep --page raw --csv > gpurun_out/r02r_c2_k_shade.raw.csv 2>/dev/null
ncu -i /tmp/s.ncu-rep --page source --csv > gpurun_out/r02r_c2_k_shade.source.csv 2>/dev/null
