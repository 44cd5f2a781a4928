#!/bin/bash
# eight GPUs: the torchrun bench line (weak headline, strong block, one-call multi-device form) and the multi-device tests
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/b9_n8.json 2> gpurun_out/b9_n8.err; tail -c 400 gpurun_out/b9_n8.err
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t9.log 2>&1; tail -3 gpurun_out/t9.log
