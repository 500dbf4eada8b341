"""profiles/ncu_summary.json: DRAM bytes per launch of the dominant Chamfer kernel, read by bench.py for
`roofline.traffic`.  Stamped with the kernel name, the capture and the hash of the kernel sources it was taken
from -- bench.py refuses the entry when csrc/chamfer.cu, chamfer_sweep.cu or tc_common.cuh has changed since.
    python tools/ncu_traffic.py <workload> <capture name under gpurun_out/> [<workload> <capture> ...]"""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sha = hashlib.sha256(b"".join(open(os.path.join(ROOT, "pytorch_points_b200", "csrc", f), "rb").read()
                              for f in ("chamfer.cu", "chamfer_sweep.cu", "tc_common.cuh"))).hexdigest()[:16]
path = os.path.join(ROOT, "profiles", "ncu_summary.json")
out = {}
args = sys.argv[1:]
for workload, cap in zip(args[0::2], args[1::2]):
    rep = os.path.join(ROOT, "gpurun_out", cap + ".ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, r = rows[0], rows[1], rows[2]
    d = {h[i]: (r[i], u[i]) for i in range(len(h))}
    def nbytes(k):
        v, unit = d[k]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    out[workload] = {"kernel": d["Kernel Name"][0], "capture": "gpurun_out/%s.ncu-rep (ncu --set full --clock-control none)" % cap,
                     "chamfer_cu_sha16": sha,
                     "chamfer_fwd_dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
                     "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
                     "gpu_time_us": d["gpu__time_duration.sum"][0] + " " + d["gpu__time_duration.sum"][1]}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
