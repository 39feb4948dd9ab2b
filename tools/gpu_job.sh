python -m pytest tests -x -q -m gpu -k "posteriors_error or family_kernel_vs" > gpurun_out/t35.log 2>&1; echo "rc=$?" >> gpurun_out/t35.log
for cfg in 0 1 2; do BPPGPU_PRUNE_CFG=$cfg python bench.py --workload protein_g4_500x200k --steps 10 --warmup 3 --no-cpu 2>&1 >/dev/null | grep "timed region" | sed "s/^/cfg=$cfg /" >> gpurun_out/sweep_prune_cfg.log; done
