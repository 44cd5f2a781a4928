#!/bin/bash
# two GPUs: multi-device tests, the torchrun bench line (both arms)
python -m pytest tests/test_gpu_multi.py tests/test_gpu_grad.py -x -q > gpurun_out/t7.log 2>&1; tail -3 gpurun_out/t7.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/b7_n2.json 2> gpurun_out/b7_n2.err; tail -c 600 gpurun_out/b7_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/b7_n2_ref.json 2> gpurun_out/b7_n2_ref.err; tail -c 300 gpurun_out/b7_n2_ref.err
