#!/bin/bash
# side-stream overlap sweep (run under gpurun)
for ss in 1 2 4; do for l2 in 17 33 66; do for hb in 2 4; do for fb in 2 4 8; do
  v=$(BV_SIDE_STREAMS=$ss BV_HIST_BPS=$hb BV_FINAL_BPS=$fb BV_L2_CHUNK_MB=$l2 python bench.py --steps 40 --warmup 3 --no-side --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']))")
  echo "side=$ss l2_mb=$l2 hist_bps=$hb final_bps=$fb -> $v"
done; done; done; done
