#!/bin/bash
# Steady-state DRAM traffic of one full 16-frame step with side streams ON (run under gpurun).
#  1. range replay: the whole step between cudaProfilerStart/Stop is ONE measured range, its kernels stay concurrent and
#     the caches are not flushed -- what the step really moves through HBM;
#  2. kernel replay with --cache-control none: per-kernel instruction / pipe counts (kernels serialised, caches kept).
set -x
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,gpu__time_duration.sum
for W in c2 fused; do
  python tools/traffic_run.py $W > gpurun_out/r02_traffic_plain_$W.log 2>&1 && \
  ncu --replay-mode range --cache-control none --clock-control none --metrics $M \
      --csv --log-file gpurun_out/r02_traffic_${W}_range.csv python tools/traffic_run.py $W > gpurun_out/r02_traffic_${W}_range.out 2>&1
  tail -3 gpurun_out/r02_traffic_${W}_range.out
  head -c 1200 gpurun_out/r02_traffic_${W}_range.csv
done
if [ -n "$PERKERNEL" ]; then
  P=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum
  for W in c2 fused; do
    ncu --cache-control none --clock-control none --profile-from-start off --metrics $P \
        --csv --log-file gpurun_out/r02_perkernel_$W.csv python tools/traffic_run.py $W > gpurun_out/r02_perkernel_$W.out 2>&1
  done
fi
