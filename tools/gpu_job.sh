python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/t10.log
python bench.py --workload codon_200x100k --steps 5 --warmup 3 --no-cpu > gpurun_out/b_codon2.json 2> gpurun_out/b_codon2.err
python bench.py --workload protein_g4_500x200k_d2 --steps 3 --warmup 3 --no-cpu > gpurun_out/b_prot2.json 2> gpurun_out/b_prot2.err
