#!/bin/bash
# usage: tools/ncu_summary.sh <report.ncu-rep> <kernel-substring> [top]  -> metrics + source hot spots
set -e
REP=$1; KERN=$2; TOP=${3:-40}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
(cd $TMP && cuobjdump -xelf all $ROOT/physics_reinforcement_learning_environment_b200/libprl.so > /dev/null)
ncu -i $REP --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','smsp__inst_executed_op_shared_ld.sum']
print('metric,unit,value')
for h,u,v in zip(hdr,units,vals):
    if h in want or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(v)>0.05): print(f'{h},{u},{v}')
"
ncu -i $REP --page source --csv 2>/dev/null > $TMP/src.csv
python $ROOT/tools/ncu_lines.py $TMP/src.csv $KERN $TMP/prl_api.sm_100a.cubin $TOP
rm -rf $TMP
