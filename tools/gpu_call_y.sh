#!/bin/bash
# round 2, call Y: LZX P1 with the slow LENGTH load consumed inside its branch (no long-scoreboard wait on the add every match step runs)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_chm_frontend.py -m gpu -q -k "lzx or golden or corrupt or unaligned or mixed or interval" ) > gpurun_out/y_pytest_lzx.log 2>&1; tail -4 gpurun_out/y_pytest_lzx.log
( QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 3 ) > gpurun_out/y_qb_lzx.log 2>&1; echo "lzx: $(grep 'stage_timing=True' gpurun_out/y_qb_lzx.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/y_qb_lzx.log) $(grep -o 'best [0-9.]* ms = [0-9.]* GB/s' gpurun_out/y_qb_lzx.log)"
