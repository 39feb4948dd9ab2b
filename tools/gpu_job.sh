python -m pytest tests -x -q -m gpu -k "chromosome or points" > gpurun_out/t37.log 2>&1; echo "rc=$?" >> gpurun_out/t37.log
for th in 256 512 1024; do BPPGPU_POINTS_THREADS=$th python bench.py --workload chromosome_500x4096pts --points 256 --steps 2 --warmup 1 --no-cpu 2>&1 >/dev/null | grep "timed region" | sed "s/^/threads=$th /" >> gpurun_out/sweep_points_threads.log; done
