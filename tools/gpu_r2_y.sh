# r02-y (1 GPU): compute-sanitizer memcheck over small renders of every kernel family
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r02y_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -12 gpurun_out/r02y_memcheck.log | cut -c1-200
