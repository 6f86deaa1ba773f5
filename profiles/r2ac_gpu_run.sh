#!/bin/bash
# round 2, call aa: ncu --set full of the alanine-dipeptide kernels (what bounds the fp32 SIMT path?)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:ad2_ -s 2 -c 2 -o /tmp/r2ac_ad2 -f python profiles/run_ad2.py 296 > gpurun_out/r2ac_ncu.log 2>&1; tail -1 gpurun_out/r2ac_ncu.log
python profiles/ncu_summary.py /tmp/r2ac_ad2.ncu-rep > gpurun_out/r2ac_ncu_ad2.txt
ncu -i /tmp/r2ac_ad2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:60])
    for k in h:
        if any(s in k for s in ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared','smsp__inst_executed_pipe_lsu','sm__inst_executed_pipe_alu','smsp__thread_inst_executed_per_inst_executed','sm__inst_executed.avg.per_cycle_elapsed','l1tex__lsu_writeback_active','l1tex__data_pipe_lsu_wavefronts.sum.pct')):
            print('  ', k, r[h.index(k)])
" > gpurun_out/r2ac_ncu_ad2_extra.txt
