python bench.py --steps 3 --warmup 3 --workload protein_g4_500x200k_d2 --no-cpu > gpurun_out/inv1_prot.json 2> gpurun_out/inv1_prot.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 --workload protein_g4_500x200k_d2 --scaling strong --no-cpu > gpurun_out/inv2_prot.json 2> gpurun_out/inv2_prot.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/inv1_dna.json 2> gpurun_out/inv1_dna.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong --no-cpu > gpurun_out/inv2_dna.json 2> gpurun_out/inv2_dna.err
