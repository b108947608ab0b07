#!/bin/bash
# round 2: BASELINE configs[3] and configs[4] at their full size - 1 M units over 8 B200s (131 072 units per GPU) - one process per GPU,
# launched like the driver launches bench.py; plus the host's concurrent PCIe ceiling at 8 GPUs (tools/pcie_probe.py)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m_topo8.txt 2>&1; nproc >> gpurun_out/m_topo8.txt; free -g >> gpurun_out/m_topo8.txt
run() { cfg=$1; shift; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config $cfg --steps 3 --warmup 3 --e2e-inflight 1 "$@" ) > gpurun_out/m_bench8_cfg$cfg.log 2>&1; grep "^{" gpurun_out/m_bench8_cfg$cfg.log | cut -c1-260; tail -3 gpurun_out/m_bench8_cfg$cfg.log | grep real; }
run 4
run 5
( timeout 300 python tools/pcie_probe.py 8 ) > gpurun_out/m_pcie_probe8.log 2>&1; tail -3 gpurun_out/m_pcie_probe8.log
