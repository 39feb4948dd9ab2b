python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/t5.log
BPPGPU_WALK4_PIPE=1 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench4_pipe.json 2> gpurun_out/bench4_pipe.err
