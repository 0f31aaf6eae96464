#!/bin/bash
# run on the GPU box: synthetic frames from the test generator, then the micro-benchmarks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python - <<'PY'
import numpy as np
from oracle import synth
np.stack([synth.gen_underwater(1242, 2208, 10 + i) for i in range(8)]).tofile('/tmp/frames.raw')
PY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ubench_hist.log
tools/ubench/hist_bench /tmp/frames.raw ${UBENCH_ARGS} >> gpurun_out/ubench_hist.log 2>&1
tail -40 gpurun_out/ubench_hist.log
if [ -n "$UBENCH_NCU" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"$UBENCH_NCU" -o gpurun_out/ubench_prof -f \
      tools/ubench/hist_bench /tmp/frames.raw ncu > gpurun_out/ubench_ncu.log 2>&1
  tail -5 gpurun_out/ubench_ncu.log
fi
