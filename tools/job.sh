set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s9_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s9_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s9_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/s9_smoke.log
timeout 300 python bench.py --workload chromosome_500x4096pts --well-conditioned --steps 5 --warmup 3 > gpurun_out/s9_bench_chr_wc.json 2> gpurun_out/s9_bench_chr_wc.err
BPPGPU_COPY_THREADS=4 timeout 300 python bench.py --workload chromosome_500x4096pts --well-conditioned --steps 5 --warmup 3 --no-cpu > gpurun_out/s9_bench_chr_wc_t4.json 2> gpurun_out/s9_bench_chr_wc_t4.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/s9_bench_dna.json 2> gpurun_out/s9_bench_dna.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s9_bench_ref.json 2> gpurun_out/s9_bench_ref.err
du -sh gpurun_out
