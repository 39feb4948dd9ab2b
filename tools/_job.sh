set -x
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab_streamed or level_batched or factored_points" > gpurun_out/s15_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/s15_memcheck.log
tail -c 3000 gpurun_out/s15_memcheck.log > gpurun_out/s15_memcheck_tail.log
