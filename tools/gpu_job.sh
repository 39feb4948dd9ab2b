python -m pytest tests -x -q -m gpu -k "posteriors or shim" > gpurun_out/t34.log 2>&1; echo "rc=$?" >> gpurun_out/t34.log
