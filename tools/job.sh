set -x
timeout 600 ncu -k regex:"walk4c_kernel" --set full --clock-control none -c 1 -o /tmp/w4c python bench.py --profile > gpurun_out/s11_ncu_w4c.log 2>&1
python tools/ncu_summary.py /tmp/w4c.ncu-rep gpurun_out/s11_walk4c_summary.csv
timeout 600 ncu -k regex:"dmma_prune_level" --set full --clock-control none --launch-skip 20 -c 6 -o /tmp/p64 python bench.py --workload codon_200x100k --profile > gpurun_out/s11_ncu_p64.log 2>&1
python tools/ncu_summary.py /tmp/p64.ncu-rep gpurun_out/s11_prune_level64_summary.csv
timeout 600 ncu -k regex:"dmma_prune_level" --set full --clock-control none --launch-skip 30 -c 6 -o /tmp/p20 python bench.py --workload protein_g4_500x200k --profile > gpurun_out/s11_ncu_p20.log 2>&1
python tools/ncu_summary.py /tmp/p20.ncu-rep gpurun_out/s11_prune_level20_summary.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s11_launches_protval.csv python bench.py --workload protein_g4_500x200k --profile > gpurun_out/s11_ncu_protval.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/s11_launches_dna.csv python bench.py --profile > gpurun_out/s11_ncu_dna.log 2>&1
du -sh gpurun_out
