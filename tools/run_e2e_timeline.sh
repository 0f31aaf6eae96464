python tools/e2e_timeline.py 2> gpurun_out/r02_e2e_timeline.log
for mb in 8 16 17 25 33; do echo "BV_HOST_CHUNK_MB=$mb" >> gpurun_out/r02_e2e_timeline.log; BV_HOST_CHUNK_MB=$mb python tools/e2e_timeline.py 2>&1 | grep stage_host >> gpurun_out/r02_e2e_timeline.log; done
python tools/e2e_timeline.py --timeline 2>&1 | tail -16 >> gpurun_out/r02_e2e_timeline.log
cat gpurun_out/r02_e2e_timeline.log
