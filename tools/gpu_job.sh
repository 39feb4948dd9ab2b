python -m pytest tests -m gpu -x -q -k "three_classes" > gpurun_out/t39.log 2>&1; echo "rc=$?" >> gpurun_out/t39.log
