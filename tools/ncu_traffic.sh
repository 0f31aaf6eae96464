set -x
python tools/traffic_run.py c2 > gpurun_out/r02_traffic_plain_c2.log 2>&1 && \
ncu --replay-mode range --cache-control none --clock-control none --profile-from-start off \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,gpu__time_duration.sum \
    --csv --log-file gpurun_out/r02_traffic_c2_range.csv python tools/traffic_run.py c2 > gpurun_out/r02_traffic_c2_range.out 2>&1
tail -3 gpurun_out/r02_traffic_c2_range.out
python tools/traffic_run.py fused > gpurun_out/r02_traffic_plain_fused.log 2>&1 && \
ncu --replay-mode range --cache-control none --clock-control none --profile-from-start off \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,gpu__time_duration.sum \
    --csv --log-file gpurun_out/r02_traffic_fused_range.csv python tools/traffic_run.py fused > gpurun_out/r02_traffic_fused_range.out 2>&1
tail -3 gpurun_out/r02_traffic_fused_range.out
# per-kernel (kernel replay, caches NOT flushed between kernels): executed instructions and DRAM bytes of every launch of the profiled step
ncu --cache-control none --clock-control none --profile-from-start off \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum \
    --csv --log-file gpurun_out/r02_perkernel_c2.csv python tools/traffic_run.py c2 > gpurun_out/r02_perkernel_c2.out 2>&1
ncu --cache-control none --clock-control none --profile-from-start off \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum \
    --csv --log-file gpurun_out/r02_perkernel_fused.csv python tools/traffic_run.py fused > gpurun_out/r02_perkernel_fused.out 2>&1
head -c 1500 gpurun_out/r02_traffic_c2_range.csv
