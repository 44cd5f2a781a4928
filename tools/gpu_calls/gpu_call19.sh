#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/b19_n8.json 2> gpurun_out/b19_n8.err; tail -c 300 gpurun_out/b19_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/b19_n4.json 2> gpurun_out/b19_n4.err; tail -c 300 gpurun_out/b19_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b19_n2.json 2> gpurun_out/b19_n2.err; tail -c 300 gpurun_out/b19_n2.err
