python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/t11.log
