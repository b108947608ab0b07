#!/bin/bash
# usage: tools/ncu_summary.sh gpurun_out/prof_X.ncu-rep > profiles/X.txt
# Text summary of one ncu --set full capture: key raw metrics + hottest CUDA source lines.
rep="$1"
echo "# ncu summary of $rep"
ncu -i "$rep" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; units=rows[1]; vals=rows[2]
want=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__shared_mem_per_block_static','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.max']
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(f'{w:70s} {vals[i]} {units[i]}')
"
echo
echo "# hottest source lines (share of warp-stall samples / of executed warp instructions / active threads per instruction)"
ncu -i "$rep" --page source --print-source cuda,sass --csv 2>/dev/null | python "$(dirname "$0")/ncu_lines.py" 25
