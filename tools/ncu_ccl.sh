#!/bin/bash
# ncu --set full of the labelling kernels on the C5 workload (4 frames 3840x2160 per launch), run under gpurun after the
# same command exited 0 without ncu.
set -x
python tools/profile_run.py c5 2 > gpurun_out/r02_ccl_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"ccl_final|ccl_merge|ccl_count" -s 3 -c 3 \
    -o gpurun_out/r02_ccl_full -f python tools/profile_run.py c5 2 > gpurun_out/r02_ccl_ncu.out 2>&1
tail -3 gpurun_out/r02_ccl_ncu.out
ls -la gpurun_out/r02_ccl_full.ncu-rep
