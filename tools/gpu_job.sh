python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t7.log
for pf in 1 0; do for pt in 1 2 4; do BPPGPU_WALK4_PREFETCH=$pf BPPGPU_WALK4_PT=$pt python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench7_pf${pf}_pt$pt.json 2> gpurun_out/bench7.err; done; done
