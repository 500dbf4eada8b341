#!/bin/bash
# Run on the GPU box (under gpurun): one `ncu --set full` capture per hot-path kernel plus the
# launch list of the default bench.  Reports land in gpurun_out/; tools/ncu_summarize.py turns
# them into profiles/*.json here.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
# variant 0 = default choice, blocks-per-SM 24 = default split heuristic
timeout 300 $NCU -k regex:chamfer_fwd_kernel -c 1 -s 2 -o gpurun_out/prof_ch2500 python tools/prof_chamfer.py 32 2500 0 24 > gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:chamfer_fwd_kernel -c 1 -s 2 -o gpurun_out/prof_ch8192 python tools/prof_chamfer.py 32 8192 0 24 >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:chamfer_finalize -c 1 -s 2 -o gpurun_out/prof_fin2500 python tools/prof_chamfer.py 32 2500 0 24 >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:chamfer_bwd -c 2 -s 4 -o gpurun_out/prof_bwd2500 python tools/prof_chamfer.py 32 2500 0 24 >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:fps_cluster -c 1 -s 1 -o gpurun_out/prof_fps python tools/prof_all.py fps >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:ball_query -c 1 -s 1 -o gpurun_out/prof_bq python tools/prof_all.py bq >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:query_group_kernel -c 1 -s 1 -o gpurun_out/prof_qg python tools/prof_all.py qg >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:knn_sweep -c 1 -s 1 -o gpurun_out/prof_knn8k python tools/prof_all.py knn 32 8192 >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:knn_sweep -c 1 -s 1 -o gpurun_out/prof_knn131k python tools/prof_all.py knn 4 131072 >> gpurun_out/cap.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c "==PROF== Report" gpurun_out/cap.log
