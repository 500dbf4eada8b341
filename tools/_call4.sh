mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_parity_gpu.py -q -x -k "all_variants" 2>&1 | tail -3
timeout 100 python tools/ch_time.py 2,22,5,25,1,21 32x2500,32x2048,32x8192,32x1024,8x2500) 2>&1 | tee gpurun_out/call_stageq.log
