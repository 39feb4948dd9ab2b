python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t15.log
