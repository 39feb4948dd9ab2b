set -x
K='regex:dmma_|pt_|walk4|generic_|finalize_|tiptab|_pack|pack_|transpose_codes'
timeout 300 ncu -k "$K" --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s12_launches_codon.csv python bench.py --workload codon_200x100k --profile > gpurun_out/s12_ncu_codon.log 2>&1
timeout 300 ncu -k "$K" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s12_launches_protval.csv python bench.py --workload protein_g4_500x200k --profile > gpurun_out/s12_ncu_protval.log 2>&1
timeout 300 ncu -k "$K" --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s12_launches_dna.csv python bench.py --profile > gpurun_out/s12_ncu_dna.log 2>&1
du -sh gpurun_out
