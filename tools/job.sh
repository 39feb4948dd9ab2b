set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s13_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s13_tests.log
timeout 300 python bench.py --workload chromosome_500x4096pts --well-conditioned --steps 5 --warmup 3 > gpurun_out/s13_bench_chr_wc.json 2> gpurun_out/s13_bench_chr_wc.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s13_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/s13_smoke.log
