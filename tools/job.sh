set -x
timeout 300 python bench.py --workload chromosome_500x4096pts --well-conditioned --steps 5 --warmup 3 > gpurun_out/s4_bench_chr_wc.json 2> gpurun_out/s4_bench_chr_wc.err
BPPGPU_CHR_CHAIN_CTAS=1 timeout 300 python bench.py --workload chromosome_500x4096pts --well-conditioned --steps 5 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "timed region" > gpurun_out/s4_chr_chain_ctas1.txt
timeout 900 python bench.py --workload chromosome_500x4096pts --steps 2 --warmup 1 > gpurun_out/s4_bench_chr_all.json 2> gpurun_out/s4_bench_chr_all.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s4_tests.log
for lb in 1 0; do
BPPGPU_LEVEL_BATCH=$lb timeout 300 python bench.py --workload protein_g4_500x200k_d2 --steps 5 --warmup 3 --no-cpu > gpurun_out/s4_bench_prot_lb$lb.json 2> gpurun_out/s4_bench_prot_lb$lb.err
BPPGPU_LEVEL_BATCH=$lb timeout 300 python bench.py --workload protein_g4_500x200k --steps 10 --warmup 3 --no-cpu > gpurun_out/s4_bench_protval_lb$lb.json 2> gpurun_out/s4_bench_protval_lb$lb.err
BPPGPU_LEVEL_BATCH=$lb timeout 300 python bench.py --workload codon_200x100k --steps 10 --warmup 3 --no-cpu > gpurun_out/s4_bench_codon_lb$lb.json 2> gpurun_out/s4_bench_codon_lb$lb.err
done
timeout 600 ncu -k regex:"chr_chain_slab" --set full --clock-control none --import-source on -c 1 -o /tmp/chain python bench.py --workload chromosome_500x4096pts --well-conditioned --points 1024 --profile > gpurun_out/s4_ncu_chr.log 2>&1
python tools/ncu_summary.py /tmp/chain.ncu-rep gpurun_out/s4_chr_chain_slab_summary.csv
python tools/ncu_source_hot.py /tmp/chain.ncu-rep chr_chain 0 60 > gpurun_out/s4_chr_chain_slab_hot.txt 2>&1
timeout 600 ncu -k regex:"chr_level_slab" --set full --clock-control none --import-source on --launch-skip 1 -c 1 -o /tmp/level python bench.py --workload chromosome_500x4096pts --well-conditioned --points 1024 --profile >> gpurun_out/s4_ncu_chr.log 2>&1
python tools/ncu_summary.py /tmp/level.ncu-rep gpurun_out/s4_chr_level_slab_summary.csv
python tools/ncu_source_hot.py /tmp/level.ncu-rep chr_level 0 60 > gpurun_out/s4_chr_level_slab_hot.txt 2>&1
timeout 600 ncu -k regex:"pt_series" --set full --clock-control none --import-source on -c 1 -o /tmp/series python bench.py --workload chromosome_500x4096pts --points 64 --profile > gpurun_out/s4_ncu_series.log 2>&1
python tools/ncu_summary.py /tmp/series.ncu-rep gpurun_out/s4_pt_series_summary.csv
python tools/ncu_source_hot.py /tmp/series.ncu-rep pt_series 0 40 > gpurun_out/s4_pt_series_hot.txt 2>&1
du -sh gpurun_out
