#!/bin/bash
# 2 GPUs: sharded tests + bench plumbing (NCCL through the engine, no torch)
O=gpurun_out/m2; mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rs -x > $O/sharded_tests.log 2>&1; echo "sharded rc=$?"
tail -12 $O/sharded_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 1 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"
tail -c 1500 $O/bench_n2.json; tail -5 $O/bench_n2.err
