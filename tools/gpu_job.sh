python -m pytest tests -m gpu -x -q > gpurun_out/t31.log 2>&1; echo "rc=$?" >> gpurun_out/t31.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "rc=$?" >> gpurun_out/smoke2.log
timeout 600 python bench.py > gpurun_out/f_default.json 2> gpurun_out/f_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_reference.json 2> gpurun_out/f_reference.err
timeout 600 python bench.py --workload protein_g4_500x200k_d2 --steps 5 --warmup 3 > gpurun_out/f_prot_d2.json 2> gpurun_out/f_prot_d2.err
timeout 600 python bench.py --workload protein_g4_500x200k --steps 10 --warmup 3 --no-cpu > gpurun_out/f_prot_val.json 2> gpurun_out/f_prot_val.err
timeout 600 python bench.py --workload codon_200x100k --steps 10 --warmup 3 > gpurun_out/f_codon.json 2> gpurun_out/f_codon.err
timeout 900 python bench.py --workload chromosome_500x4096pts --steps 2 --warmup 1 --no-cpu > gpurun_out/f_chr4096.json 2> gpurun_out/f_chr4096.err
