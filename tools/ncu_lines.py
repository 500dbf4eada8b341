"""Per-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, os, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg = {}
fname, h = "?", None
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        h = r; ia = h.index("Instructions Executed"); ismp = h.index("# Samples"); continue
    if h is None or not r[0].strip().isdigit() or len(r) <= ia:
        continue
    a = agg.setdefault((fname, int(r[0]), r[1].strip()), [0, 0])
    num = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
    a[0] += num(r[ia]); a[1] += num(r[ismp])
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total inst", tot, "samples", tots)
for (f, ln, src), v in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-22s %5d  inst %5.1f%%  smp %5.1f%%  %s" % (f[:22], ln, 100 * v[0] / max(tot, 1), 100 * v[1] / max(tots, 1), src[:100]))
