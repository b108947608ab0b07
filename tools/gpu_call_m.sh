#!/bin/bash
# round 2, call M: Quantum with the hot / cold split of the frequency tables (448 lanes per SM): gpu tier, Quantum line, mixed line, ncu of k_p1_qtm; LZX headline with the deferred LENGTH load
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/m_pytest_gpu.log 2>&1; tail -4 gpurun_out/m_pytest_gpu.log
( timeout 600 python bench.py --config 6 --steps 3 --e2e-inflight 1 ) > gpurun_out/m_bench_cfg6.log 2>&1; grep "^{" gpurun_out/m_bench_cfg6.log | cut -c1-200
( timeout 900 python bench.py --config 5 --steps 3 --e2e-inflight 1 ) > gpurun_out/m_bench_cfg5.log 2>&1; grep "^{" gpurun_out/m_bench_cfg5.log | cut -c1-200
( QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 4 ) > gpurun_out/m_qb_lzx.log 2>&1; echo "lzx: $(grep 'stage_timing=True' gpurun_out/m_qb_lzx.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/m_qb_lzx.log)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_p1_qtm" -c 1 -f -o gpurun_out/m_prof_p1qtm python tools/quickbench.py 2 65536 1 > gpurun_out/m_ncu_qtm.log 2>&1; tail -1 gpurun_out/m_ncu_qtm.log
