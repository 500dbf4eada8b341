"""Group the SASS rows of an `ncu --page source --csv` dump into runs of equal execution count and
print the heaviest ones (instruction share, sample share, top opcodes, top stall reasons)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; data = rows[2:]
ia = h.index("Instructions Executed"); isrc = h.index("Source"); ismp = h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[ismp]) for r in data)
print("total inst", tot, "samples", tots, "rows", len(data))
segs = []; cur = None
for i, r in enumerate(data):
    n = int(r[ia])
    if cur and abs(n - cur['n']) <= 0.03 * max(n, cur['n'], 1):
        cur['end'] = i; cur['sum'] += n; cur['smp'] += int(r[ismp])
    else:
        cur = {'start': i, 'end': i, 'n': n, 'sum': n, 'smp': int(r[ismp])}; segs.append(cur)
key = 'smp' if len(sys.argv) > 2 and sys.argv[2] == 'samples' else 'sum'
for s in sorted(segs, key=lambda s: -s[key])[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
    ops = {}; st = {}
    for r in data[s['start']:s['end'] + 1]:
        t = r[isrc].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op] = ops.get(op, 0) + 1
        for c in stall_cols:
            st[h[c]] = st.get(h[c], 0) + int(r[c] or 0)
    top = sorted(ops.items(), key=lambda x: -x[1])[:6]
    tst = sorted(st.items(), key=lambda x: -x[1])[:3]
    print("rows %4d-%4d len %3d exec %9d inst%% %4.1f smp%% %4.1f :: %s :: %s" % (
        s['start'], s['end'], s['end'] - s['start'] + 1, s['n'], 100 * s['sum'] / tot, 100 * s['smp'] / tots, top, tst))
