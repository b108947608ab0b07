#!/bin/bash
# round 2, call V: Quantum P1 with branch-free model constants, the select-tree scan8 and the register rescale of the selector
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "quantum or mixed or golden or corrupt or unaligned" ) > gpurun_out/v_pytest_qtm.log 2>&1; tail -4 gpurun_out/v_pytest_qtm.log
( time MSGPU_QTM_CONV=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "quantum or mixed or golden or corrupt" ) > gpurun_out/v_pytest_qtm_loop.log 2>&1; tail -4 gpurun_out/v_pytest_qtm_loop.log
( timeout 600 python tools/qtm_ab.py 65536 ) > gpurun_out/v_qtm_ab.log 2>&1; cat gpurun_out/v_qtm_ab.log | tail -6
