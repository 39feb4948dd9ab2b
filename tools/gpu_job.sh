python -m pytest tests/test_gpu_parity.py -m gpu -q -k "points or chromosome or pt_batch" 2>&1 | tail -40 > gpurun_out/t13.log
