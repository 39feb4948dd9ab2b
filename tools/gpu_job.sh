for pt in 3; do BPPGPU_WALK4_PT=$pt python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "timed region" | sed "s/^/pt=$pt /" >> gpurun_out/sweep_pt3.log; done
