#!/bin/bash
python tools/pcie_probe.py
for mb in 4 8 16 32 64; do
  v=$(BV_HOST_CHUNK_MB=$mb python bench.py --steps 20 --warmup 3 --no-side --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['e2e']['value']))")
  echo "host_chunk_mb=$mb -> e2e $v fps"
done
