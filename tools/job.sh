set -x
timeout 400 python -m pytest tests -m gpu -x -q -k "chr or chromosome or points or batched" > gpurun_out/s3_tests_chr.log 2>&1; echo "rc=$?" >> gpurun_out/s3_tests_chr.log
timeout 300 python bench.py --workload chromosome_500x4096pts --steps 5 --warmup 3 --no-cpu > gpurun_out/s3_bench_chr.json 2> gpurun_out/s3_bench_chr.err
for st in 2 6 12; do BPPGPU_CHR_NST_CHAIN=$st timeout 300 python bench.py --workload chromosome_500x4096pts --steps 5 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "timed region" > gpurun_out/s3_chr_chain_nst$st.txt; done
BPPGPU_CHR_NST_LEVEL=2 timeout 300 python bench.py --workload chromosome_500x4096pts --steps 5 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "timed region" > gpurun_out/s3_chr_level_nst2.txt
timeout 300 python bench.py --workload protein_g4_500x200k_d2 --steps 5 --warmup 3 --no-cpu > gpurun_out/s3_bench_prot.json 2> gpurun_out/s3_bench_prot.err
BPPGPU_PDL=0 timeout 300 python bench.py --workload protein_g4_500x200k_d2 --steps 5 --warmup 3 --no-cpu > gpurun_out/s3_bench_prot_nopdl.json 2> gpurun_out/s3_bench_prot_nopdl.err
timeout 300 python bench.py --workload codon_200x100k --steps 10 --warmup 3 --no-cpu > gpurun_out/s3_bench_codon.json 2> gpurun_out/s3_bench_codon.err
BPPGPU_PDL=0 timeout 300 python bench.py --workload codon_200x100k --steps 10 --warmup 3 --no-cpu > gpurun_out/s3_bench_codon_nopdl.json 2> gpurun_out/s3_bench_codon_nopdl.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3_tests.log
