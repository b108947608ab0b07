#!/bin/bash
# round 2, call U: ncu capture of the Quantum P1 kernel after the convergence work (converged scans are the default now)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_p1_qtm" -c 1 -f -o gpurun_out/u_prof_p1qtm python tools/quickbench.py 2 65536 1 > gpurun_out/u_ncu_qtm.log 2>&1; tail -2 gpurun_out/u_ncu_qtm.log
