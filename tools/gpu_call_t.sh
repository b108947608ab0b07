#!/bin/bash
# round 2, call T: Quantum P1 with one instruction stream per model symbol (step() without the four-way switch, merged
# renormalisation, branch-free fetch bookkeeping) and, as MSGPU_QTM_CONV=1, the eight-wide converged scans: parity + A/B
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "quantum or mixed or golden or corrupt or unaligned" ) > gpurun_out/t_pytest_qtm.log 2>&1; tail -4 gpurun_out/t_pytest_qtm.log
( time MSGPU_QTM_CONV=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "quantum or mixed or golden or corrupt or unaligned" ) > gpurun_out/t_pytest_qtm_conv.log 2>&1; tail -4 gpurun_out/t_pytest_qtm_conv.log
( timeout 600 python tools/qtm_ab.py 65536 ) > gpurun_out/t_qtm_ab.log 2>&1; cat gpurun_out/t_qtm_ab.log | tail -6
