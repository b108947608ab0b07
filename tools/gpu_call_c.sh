#!/bin/bash
# round 2, call C: the gpu tier with the new contract tests, then bench.py over every config on one GPU, the reference arm, the PCIe probe
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x ) > gpurun_out/c_pytest_gpu.log 2>&1; tail -5 gpurun_out/c_pytest_gpu.log
for c in 3 4 5 6 2 1; do
  ( time timeout 600 python bench.py --config $c --steps 5 ) > gpurun_out/c_bench_cfg$c.log 2>&1; grep "^{" gpurun_out/c_bench_cfg$c.log | cut -c1-250; grep -v "^{" gpurun_out/c_bench_cfg$c.log | tail -4
done
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/c_bench_reference.log 2>&1; grep "^{" gpurun_out/c_bench_reference.log | cut -c1-200
timeout 200 python tools/pcie_probe.py > gpurun_out/c_pcie_probe.log 2>&1; cat gpurun_out/c_pcie_probe.log
nvidia-smi topo -m > gpurun_out/c_topo.txt 2>&1; nproc >> gpurun_out/c_topo.txt; free -g >> gpurun_out/c_topo.txt
