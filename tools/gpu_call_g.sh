#!/bin/bash
# round 2, call G: gpu tier (drop-in LZX exact decode, Quantum 224 lanes + cooperative updates), Quantum line, config 4 with one / three internal streams, mixed line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/g_pytest_gpu.log 2>&1; tail -4 gpurun_out/g_pytest_gpu.log
( timeout 600 python bench.py --config 6 --steps 3 --cpu-sample 256 --e2e-inflight 1 ) > gpurun_out/g_bench_cfg6.log 2>&1; grep "^{" gpurun_out/g_bench_cfg6.log | cut -c1-200
KW4="dict(window_bits=21, unit_bytes=65536, reset_interval=2, slack=8)"
( QB_STAGE=1 timeout 600 python tools/quickbench.py 3 131072 3 "$KW4" ) > gpurun_out/g_qb_cfg4_s3.log 2>&1; tail -6 gpurun_out/g_qb_cfg4_s3.log
( MSGPU_STREAMS=1 timeout 600 python tools/quickbench.py 3 131072 3 "$KW4" ) > gpurun_out/g_qb_cfg4_s1.log 2>&1; tail -2 gpurun_out/g_qb_cfg4_s1.log
( MSGPU_STREAMS=2 timeout 600 python tools/quickbench.py 3 131072 3 "$KW4" ) > gpurun_out/g_qb_cfg4_s2.log 2>&1; tail -2 gpurun_out/g_qb_cfg4_s2.log
( MSGPU_SUBWAVE=32256 timeout 600 python tools/quickbench.py 3 131072 3 "$KW4" ) > gpurun_out/g_qb_cfg4_sw.log 2>&1; tail -2 gpurun_out/g_qb_cfg4_sw.log
( timeout 900 python bench.py --config 5 --steps 3 --cpu-sample 256 --e2e-inflight 1 ) > gpurun_out/g_bench_cfg5.log 2>&1; grep "^{" gpurun_out/g_bench_cfg5.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p1_qtm" -c 1 -f -o gpurun_out/g_prof_p1qtm python tools/quickbench.py 2 16384 1 > gpurun_out/g_ncu_qtm.log 2>&1; tail -1 gpurun_out/g_ncu_qtm.log
