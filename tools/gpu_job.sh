python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/t4.log
for pt in 1 2 4; do BPPGPU_WALK4_PT=$pt python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench3_pt$pt.json 2> gpurun_out/bench3_pt$pt.err; done
