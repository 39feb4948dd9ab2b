python -m pytest tests -m gpu -x -q -k "reparametrisation or shim" > gpurun_out/t45.log 2>&1; echo "rc=$?" >> gpurun_out/t45.log
