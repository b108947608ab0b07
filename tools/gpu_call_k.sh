#!/bin/bash
# round 2, call K: the software-pipelined resolve (MSGPU_P2_PIPE=1): parity tier with it on, A/B on the LZX and MSZIP headline batches and config 4
mkdir -p gpurun_out
( MSGPU_P2_PIPE=1 timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/k_pytest_gpu_pipe.log 2>&1; tail -2 gpurun_out/k_pytest_gpu_pipe.log
for p in 0 1; do ( MSGPU_P2_PIPE=$p QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 4 ) > gpurun_out/k_qb_lzx_pipe$p.log 2>&1; echo "lzx pipe $p: $(grep 'stage_timing=True' gpurun_out/k_qb_lzx_pipe$p.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/k_qb_lzx_pipe$p.log)"; done
for p in 0 1; do ( MSGPU_P2_PIPE=$p QB_STAGE=1 timeout 300 python tools/quickbench.py 1 65536 4 ) > gpurun_out/k_qb_zip_pipe$p.log 2>&1; echo "zip pipe $p: $(grep 'stage_timing=True' gpurun_out/k_qb_zip_pipe$p.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/k_qb_zip_pipe$p.log)"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p2_resolve" -c 1 -f -o gpurun_out/k_prof_p2pipe env MSGPU_P2_PIPE=1 python tools/quickbench.py 3 65536 1 > gpurun_out/k_ncu_p2.log 2>&1; tail -1 gpurun_out/k_ncu_p2.log
