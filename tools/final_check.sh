#!/bin/bash
# Round-end check on the GPU box: KNN captures, full GPU test-suite, smoke, default bench.
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 300 $NCU -k regex:knn_sweep -c 1 -s 1 -o gpurun_out/prof_knn8k python tools/prof_all.py knn 32 8192 > gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:knn_sweep -c 1 -s 1 -o gpurun_out/prof_knn131k python tools/prof_all.py knn 4 131072 >> gpurun_out/cap.log 2>&1
timeout 300 $NCU -k regex:km_prepare_small -c 1 -s 1 -o gpurun_out/prof_knnprep8k python tools/prof_all.py knn 32 8192 >> gpurun_out/cap.log 2>&1
grep -c "==PROF== Report" gpurun_out/cap.log
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_err.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"])
x = d["extras"]
for k in ("knn_k16_B32_N8192", "knn_k16_B4_N131072", "fps_B16_N16384_m1024"):
    print(k, x[k]["ms_per_step"])
PY
