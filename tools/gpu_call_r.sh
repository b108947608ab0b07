#!/bin/bash
# round 2, call R: Quantum P1 launches of all sub-waves issued before the first Quantum resolve launch + 32 hardware launch queues:
# host-path tests, config 5 end to end, the headline with 8 and with 32 launch queues
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host or mixed or multi_device or digest" ) > gpurun_out/r_pytest_host.log 2>&1; tail -4 gpurun_out/r_pytest_host.log
( timeout 900 python bench.py --config 5 --steps 3 --e2e-inflight 1 ) > gpurun_out/r_bench_cfg5.log 2>&1; grep "^{" gpurun_out/r_bench_cfg5.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/r_bench_cfg5.log | cut -c1-300
( timeout 600 python bench.py ) > gpurun_out/r_bench_cfg3_c32.log 2>&1; grep "^{" gpurun_out/r_bench_cfg3_c32.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/r_bench_cfg3_c32.log | cut -c1-300
( CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 600 python bench.py ) > gpurun_out/r_bench_cfg3_c8.log 2>&1; grep "^{" gpurun_out/r_bench_cfg3_c8.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/r_bench_cfg3_c8.log | cut -c1-300
