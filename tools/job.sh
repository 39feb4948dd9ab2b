set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "pt_batch or series or singular or chr or chromosome or expm or points" > gpurun_out/s8_tests_series.log 2>&1; echo "rc=$?" >> gpurun_out/s8_tests_series.log
timeout 900 python bench.py --workload chromosome_500x4096pts --steps 2 --warmup 1 > gpurun_out/s8_bench_chr_all.json 2> gpurun_out/s8_bench_chr_all.err
timeout 600 ncu -k regex:"pt_series" --metrics sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 1 --csv --log-file gpurun_out/s8_pt_series_sparse.csv python bench.py --workload chromosome_500x4096pts --points 64 --profile > gpurun_out/s8_ncu_series.log 2>&1
du -sh gpurun_out
