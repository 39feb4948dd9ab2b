set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s10_bench_dna_2gpu.json 2> gpurun_out/s10_bench_dna_2gpu.err
timeout 300 python -m pytest tests/test_engine_comm.py -m gpu -x -q > gpurun_out/s10_tests_comm.log 2>&1; echo "rc=$?" >> gpurun_out/s10_tests_comm.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload protein_g4_500x200k_d2 --steps 5 --warmup 3 --no-weak > gpurun_out/s10_bench_prot_2gpu.json 2> gpurun_out/s10_bench_prot_2gpu.err
