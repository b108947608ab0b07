#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, and the development timings of the rows added late in round 1.
# Every step has its own timeout so a slow one cannot eat the others.  Output -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
( time timeout 420 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
( time timeout 300 python bench.py ) > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/bench.log
timeout 150 python tools/quickbench.py 1 65536 4 > gpurun_out/qb_mszip.log 2>&1; tail -1 gpurun_out/qb_mszip.log
timeout 100 python tools/chain_bench.py 256 4 > gpurun_out/chain_bench.log 2>&1; tail -1 gpurun_out/chain_bench.log
timeout 100 python tools/quickbench.py 3 16384 3 "dict(window_bits=17,delta=1)" > gpurun_out/qb_delta.log 2>&1; tail -1 gpurun_out/qb_delta.log
timeout 100 python tools/quickbench.py 3 16384 3 > gpurun_out/qb_lzx16k.log 2>&1; tail -1 gpurun_out/qb_lzx16k.log
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log
