python -m pytest tests/test_cpp_shim.py -x -q -m gpu > gpurun_out/t43.log 2>&1; echo "rc=$?" >> gpurun_out/t43.log
