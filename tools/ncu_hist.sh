#!/bin/bash
# ncu --set full with source correlation of the two histogram passes (C2 workload, 4 frames per launch, side streams off)
set -x
BV_SIDE_STREAMS=1 python tools/profile_run.py c2 2 > gpurun_out/r02_hist_plain.log 2>&1 || exit 1
BV_SIDE_STREAMS=1 ncu --set full --clock-control none --import-source on -k regex:"hist_bgr|hist_sv" -s 4 -c 2 \
    -o gpurun_out/r02_hist_full -f python tools/profile_run.py c2 2 > gpurun_out/r02_hist_ncu.out 2>&1
tail -2 gpurun_out/r02_hist_ncu.out
