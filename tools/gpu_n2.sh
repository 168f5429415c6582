#!/bin/bash
N=${2:-2}
OUT=gpurun_out/${1:-n2}; mkdir -p $OUT
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > $OUT/pytest_multigpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_multigpu.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "bench rc=$?"; cat $OUT/bench_${N}gpu.json | cut -c1-600; tail -3 $OUT/bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 3 --batch 8 --new-tokens 128 > $OUT/bench_${N}gpu_b8.json 2> $OUT/bench_${N}gpu_b8.err; echo "bench b8 rc=$?"; cat $OUT/bench_${N}gpu_b8.json | cut -c1-400; tail -3 $OUT/bench_${N}gpu_b8.err
if [ "$N" = "8" ]; then timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 1 --warmup 3 --batch 8 --new-tokens 512 > $OUT/bench_${N}gpu_b8_t512.json 2> $OUT/bench_${N}gpu_b8_t512.err; echo "bench b8 t512 rc=$?"; cat $OUT/bench_${N}gpu_b8_t512.json | cut -c1-400; tail -3 $OUT/bench_${N}gpu_b8_t512.err; fi
