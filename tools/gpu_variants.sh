for v in v0 v2 v2f; do
  echo "== variant $v"
  TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so python tools/perf_probe.py 7 1024 8 2>&1 | grep "profiling=True"
  TGB200_LIB=$PWD/tungsten_b200/libtgb200_$v.so python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
done
