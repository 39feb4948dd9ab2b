set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s6_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s6_tests.log
timeout 300 python bench.py --workload chromosome_500x4096pts --well-conditioned --steps 5 --warmup 3 > gpurun_out/s6_bench_chr_wc.json 2> gpurun_out/s6_bench_chr_wc.err
BPPGPU_CHR_CHAIN_CTAS=1 timeout 300 python bench.py --workload chromosome_500x4096pts --well-conditioned --steps 5 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "timed region" > gpurun_out/s6_chr_chain_ctas1.txt
timeout 300 python bench.py --workload codon_200x100k --steps 10 --warmup 3 > gpurun_out/s6_bench_codon.json 2> gpurun_out/s6_bench_codon.err
timeout 300 python bench.py --workload protein_g4_500x200k --steps 10 --warmup 3 > gpurun_out/s6_bench_protval.json 2> gpurun_out/s6_bench_protval.err
timeout 300 python bench.py --workload protein_g4_500x200k_d2 --steps 5 --warmup 3 > gpurun_out/s6_bench_prot.json 2> gpurun_out/s6_bench_prot.err
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench_dna.json 2> gpurun_out/s6_bench_dna.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s6_launches_codon.csv python bench.py --workload codon_200x100k --profile > gpurun_out/s6_ncu_codon.log 2>&1
timeout 300 ncu -k regex:"chr_|pt_dmma" --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s6_launches_chr.csv python bench.py --workload chromosome_500x4096pts --well-conditioned --profile > gpurun_out/s6_ncu_chr.log 2>&1
du -sh gpurun_out
