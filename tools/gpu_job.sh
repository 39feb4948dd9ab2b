python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/t18.log
