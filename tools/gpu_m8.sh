#!/bin/bash
# 8 GPUs: every rank count of the sharded parity tests, the default bench at N = 8 and N = 4 (weak + 256-seed strong +
# synthetic 50 M population), per-phase device times of the sharded day
O=gpurun_out/m8; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_sharded.py -v -m gpu -rs > $O/sharded_tests_8gpu.log 2>&1; echo "sharded rc=$?"
tail -22 $O/sharded_tests_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench8 rc=$?"
tail -c 1300 $O/bench_n8.json
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench4 rc=$?"
tail -c 900 $O/bench_n4.json
RB_SHARD_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --workload synth50m --steps 2 --warmup 1 > $O/synth_n8.json 2> $O/synth_n8_phases.err; echo "synth8 rc=$?"
cat $O/synth_n8.json; grep "rank 0" $O/synth_n8_phases.err | tail -2
