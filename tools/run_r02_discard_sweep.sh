python -m pytest tests/test_gpu_balance_stage.py tests/test_gpu_c5.py -x -q 2>&1 | tail -4
bash tools/ncu_traffic.sh > gpurun_out/r02_ncu_traffic2.log 2>&1
grep -h '^"0"' gpurun_out/r02_traffic_c2_range.csv gpurun_out/r02_traffic_fused_range.csv | awk -F'","' '{print $11, $13}'
for mb in 17 33; do for ss in 2 4; do
  echo "chunk_mb=$mb side=$ss"; BV_L2_CHUNK_MB=$mb BV_SIDE_STREAMS=$ss python bench.py --steps 30 --no-cpu --no-side 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['sustained']['value']), d['stage_roofline']['per_kernel_ms'])"
done; done
