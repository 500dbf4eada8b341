#!/bin/bash
# Round-end check on the GPU box: full GPU test-suite, smoke, default bench, launch list and one
# full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"])
print("fused", d["fused_step"])
x = d["extras"]
for k in ("chamfer_fwd_bwd_B32_N8192", "knn_k16_B32_N8192", "knn_k16_B4_N131072", "fps_B16_N16384_m1024"):
    print(k, x[k]["ms_per_step"])
PY
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 100 $NCU -k regex:chamfer_fwd_kernel -c 1 -s 2 -o gpurun_out/prof_ch2500 python tools/prof_chamfer.py 32 2500 0 24 fused > gpurun_out/cap.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_default.csv \
    python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c "==PROF== Report" gpurun_out/cap.log
