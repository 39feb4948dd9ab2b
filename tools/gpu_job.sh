python -m pytest tests -m gpu -x -q > gpurun_out/t42.log 2>&1; echo "rc=$?" >> gpurun_out/t42.log
timeout 600 python bench.py --workload codon_200x100k --steps 10 --warmup 3 > gpurun_out/g_codon.json 2> gpurun_out/g_codon.err
python bench.py --workload chromosome_500x4096pts --points 256 --steps 2 --warmup 1 --no-cpu 2>&1 >/dev/null | grep "timed region" > gpurun_out/chr256_final.log
