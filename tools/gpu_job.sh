# last verification of the tree as committed: GPU parity suite, smoke, default bench, a short codon bench
python -m pytest tests -m gpu -x -q > gpurun_out/t44.log 2>&1; echo "rc=$?" >> gpurun_out/t44.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; echo "rc=$?" >> gpurun_out/smoke4.log
timeout 600 python bench.py > gpurun_out/h_default.json 2> gpurun_out/h_default.err
timeout 300 python bench.py --workload codon_200x100k --steps 5 --warmup 3 --no-cpu > gpurun_out/h_codon.json 2> gpurun_out/h_codon.err
