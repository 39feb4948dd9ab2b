python -m pytest tests -m gpu -x -q > gpurun_out/t38.log 2>&1; echo "rc=$?" >> gpurun_out/t38.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "rc=$?" >> gpurun_out/smoke3.log
timeout 600 python bench.py > gpurun_out/g_default.json 2> gpurun_out/g_default.err
timeout 600 python bench.py --workload protein_g4_500x200k_d2 --steps 5 --warmup 3 > gpurun_out/g_prot_d2.json 2> gpurun_out/g_prot_d2.err
timeout 600 python bench.py --workload protein_g4_500x200k --steps 10 --warmup 3 --no-cpu > gpurun_out/g_prot_val.json 2> gpurun_out/g_prot_val.err
timeout 900 python bench.py --workload chromosome_500x4096pts --steps 2 --warmup 1 --no-cpu > gpurun_out/g_chr4096.json 2> gpurun_out/g_chr4096.err
