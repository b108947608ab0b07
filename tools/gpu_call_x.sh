#!/bin/bash
# round 2, call X: bench lines of config 5 (mixed) and config 6 (Quantum) with the converged Quantum kernel and the byte-wise LZX literals
mkdir -p gpurun_out
( timeout 900 python bench.py --config 5 --steps 3 --e2e-inflight 1 ) > gpurun_out/x_bench_cfg5.log 2>&1; grep "^{" gpurun_out/x_bench_cfg5.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/x_bench_cfg5.log | cut -c1-300
( timeout 900 python bench.py --config 6 --steps 3 --e2e-inflight 1 ) > gpurun_out/x_bench_cfg6.log 2>&1; grep "^{" gpurun_out/x_bench_cfg6.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/x_bench_cfg6.log | cut -c1-300
