#!/bin/bash
# First GPU call of round 2: everything written at the end of round 1 without GPU time left.
#   gpurun --timeout 600 -- 'bash tools/r02_first_call.sh'
mkdir -p gpurun_out
(PP_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -k "all_variants or sampled_dense_edge_conv" 2>&1 | tail -5
timeout 200 python tools/r02_fold_ab.py
# staged query tiles at N = 8192 with the L2 flushed (the reason large clouds went back to variant 1)
timeout 100 python tools/ch_fused_time.py 32x8192) 2>&1 | tee gpurun_out/r02_first_call.log
