#!/bin/bash
# quick tuning sweep (run under gpurun): device-resident C2 throughput for a few knob settings
for hb in 4 8 16; do for fb in 3 4 8; do for l2 in 40 66; do
  v=$(BV_HIST_BPS=$hb BV_FINAL_BPS=$fb BV_L2_CHUNK_MB=$l2 python bench.py --steps 30 --warmup 3 --no-side --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['stage_roofline']['per_kernel_ms'], round(d['e2e']['value']))")
  echo "hist_bps=$hb final_bps=$fb l2_mb=$l2 -> $v"
done; done; done
