#!/bin/bash
# Run on the GPU box (under gpurun): `ncu --set full` captures of the Chamfer kernels at the headline workload
# (B=256, N=M=8192) and of the tensor-core k-NN kernels (B=32, N=8192, k=16), plus the launch list of a short
# default bench.  Reports land in gpurun_out/; tools/ncu_summarize.py / tools/ncu_traffic.py turn them into
# profiles/*.json here.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:cs_rowpass_tc -c 1 -s 1 -o gpurun_out/r02_tc_b256 python tools/prof_chamfer.py 256 8192 0 24 fused > gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:cs_rowpass_tc -c 1 -s 1 -o gpurun_out/r02_tc_b32 python tools/prof_chamfer.py 32 8192 0 24 fused >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:cs_finalize -c 1 -s 1 -o gpurun_out/r02_fin_b256 python tools/prof_chamfer.py 256 8192 0 24 fused >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:cs_prep -c 1 -s 1 -o gpurun_out/r02_prep_b256 python tools/prof_chamfer.py 256 8192 0 24 fused >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:chamfer_fwd_kernel -c 1 -s 1 -o gpurun_out/r02_exact_b256 python tools/prof_chamfer.py 256 8192 1 24 fused >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:kt_rowpass -c 1 -s 1 -o gpurun_out/r02_kt_rowpass python tools/knn_tc_prof.py >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:kt_select -c 1 -s 1 -o gpurun_out/r02_kt_select python tools/knn_tc_prof.py >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:kt_seed -c 1 -s 1 -o gpurun_out/r02_kt_seed python tools/knn_tc_prof.py >> gpurun_out/cap.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench_default.csv \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c "==PROF== Report" gpurun_out/cap.log
