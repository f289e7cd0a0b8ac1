"""Opcode mix (weighted by executed warp instructions) and top stall sites from an `ncu --page source --csv` export."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
kern = None
hdr = None
data = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; data[kern] = []; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and hdr and len(r) >= 6:
        data[kern].append(dict(zip(hdr, r)))
for kern, rs in data.items():
    print("==", kern[:140])
    mix = collections.Counter(); tot = 0; samples = 0
    for d in rs:
        op = d["Source"].split()
        if not op: continue
        o = op[1] if op[0].startswith("@") else op[0]
        o = o.split(".")[0]
        n = int(d["Instructions Executed"] or 0)
        mix[o] += n; tot += n; samples += int(d["# Samples"] or 0)
    print("  total warp instructions", tot, " samples", samples)
    for o, n in mix.most_common(22):
        print(f"  {o:12s} {n:12d} {100*n/tot:5.1f}%")
    top = sorted(rs, key=lambda d: -int(d["# Samples"] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for d in top:
        st = sorted(((int(d[c] or 0), c) for c in stall_cols), reverse=True)[:2]
        print(f"  {int(d['# Samples']):7d} {100*int(d['# Samples'])/max(samples,1):5.1f}%  {d['Source'].strip()[:70]:70s} {st}")
