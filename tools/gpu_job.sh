python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dmma or codon or protein or derivatives" 2>&1 | tail -5 > gpurun_out/t16.log
for rw in 1 2; do BPPGPU_DMMA_RW=$rw python bench.py --workload codon_200x100k --steps 5 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "timed region" | sed "s/^/codon rw=$rw /" >> gpurun_out/sweep_dmma2.log; done
for rw in 1 2; do BPPGPU_DMMA_RW=$rw python bench.py --workload protein_g4_500x200k_d2 --steps 2 --warmup 3 --no-cpu > gpurun_out/sweep_prot_rw$rw.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/sweep_prot_rw$rw.json')); print('protein rw=$rw', d['ms_per_step'], 'prune', d['roofline']['kernel_ms'])" >> gpurun_out/sweep_dmma2.log; done
