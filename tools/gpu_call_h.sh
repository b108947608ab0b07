#!/bin/bash
# round 2, call H: P2 owner-form pass A (A/B on LZX and MSZIP), one sub-wave for device buffers (config 4), stream per codec for mixed batches (config 5), gpu tier
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/h_pytest_gpu.log 2>&1; tail -4 gpurun_out/h_pytest_gpu.log
for o in 0 1; do ( MSGPU_P2_OWNER=$o QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 5 ) > gpurun_out/h_qb_lzx_owner$o.log 2>&1; tail -3 gpurun_out/h_qb_lzx_owner$o.log; done
for o in 0 1; do ( MSGPU_P2_OWNER=$o QB_STAGE=1 timeout 300 python tools/quickbench.py 1 65536 5 ) > gpurun_out/h_qb_zip_owner$o.log 2>&1; tail -3 gpurun_out/h_qb_zip_owner$o.log; done
KW4="dict(window_bits=21, unit_bytes=65536, reset_interval=2, slack=8)"
( QB_STAGE=1 timeout 600 python tools/quickbench.py 3 131072 3 "$KW4" ) > gpurun_out/h_qb_cfg4.log 2>&1; tail -4 gpurun_out/h_qb_cfg4.log
( timeout 900 python bench.py --config 5 --steps 3 --cpu-sample 256 --e2e-inflight 1 ) > gpurun_out/h_bench_cfg5.log 2>&1; grep "^{" gpurun_out/h_bench_cfg5.log | cut -c1-200
( MSGPU_STREAMS=1 timeout 900 python bench.py --config 5 --steps 3 --cpu-sample 256 --e2e-inflight 1 ) > gpurun_out/h_bench_cfg5_s1.log 2>&1; grep "^{" gpurun_out/h_bench_cfg5_s1.log | cut -c1-200
( timeout 600 python bench.py --steps 10 ) > gpurun_out/h_bench_cfg3.log 2>&1; grep "^{" gpurun_out/h_bench_cfg3.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p2_resolve" -c 1 -f -o gpurun_out/h_prof_p2 python tools/quickbench.py 3 65536 1 > gpurun_out/h_ncu_p2.log 2>&1; tail -1 gpurun_out/h_ncu_p2.log
