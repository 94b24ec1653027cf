#!/bin/bash
# usage: gpu_ncu_one.sh <kernel regex> <tag> [batch]
mkdir -p gpurun_out
B=${3:-8}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -f -o gpurun_out/prof_$2 python bench.py --steps 1 --warmup 1 --batch $B --no-cpu-baseline --sample-every 0 > gpurun_out/ncu_$2.log 2>&1; echo "exit $?"
ncu -i gpurun_out/prof_$2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sectors_srcunit_tex_op_read.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i,h in enumerate(hdr):
    if h in want: print(h, rows[1][i], [r[i] for r in rows[2:]])
"
