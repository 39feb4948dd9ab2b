python bench.py --workload chromosome_500x4096pts --points 256 --steps 3 --warmup 3 > gpurun_out/b_chr256.json 2> gpurun_out/b_chr256.err
python bench.py --workload chromosome_500x4096pts --steps 3 --warmup 3 > gpurun_out/b_chr.json 2> gpurun_out/b_chr.err
