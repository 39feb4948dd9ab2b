python -m pytest tests -m gpu -x -q > gpurun_out/t40.log 2>&1; echo "rc=$?" >> gpurun_out/t40.log
python bench.py --workload codon_200x100k --steps 10 --warmup 3 --no-cpu > gpurun_out/b_codon64.json 2> gpurun_out/b_codon64.err
BPPGPU_FAMILY=0 python bench.py --workload codon_200x100k --steps 10 --warmup 3 --no-cpu > gpurun_out/b_codon64_old.json 2> gpurun_out/b_codon64_old.err
