#!/bin/bash
# torchrun wrapper: rank 0 runs under ncu (only the named kernels, a handful of metrics), the other ranks run plainly.
# usage: torchrun ... --no-python profiles/ncu_rank0.sh <out.csv> <kernel-regex> <script> [args...]
out=$1; shift; kre=$1; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --clock-control none --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio \
    -k "regex:$kre" -s 6 -c 4 --csv --log-file "$out" python "$@"
else
  exec python "$@"
fi
