"""DRAM traffic per launch of the kernels in an `ncu --page raw --csv` export -> JSON {short kernel name: {launches,
dram_bytes_per_launch, time_us_per_launch, dram_gbs}} (bench.py reads profiles/r2_traffic.json for roofline.traffic)."""
import collections, csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    return v * scale
agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"^void\s+", "", r[ix["Kernel Name"]])
    name = re.sub(r"^a2v::", "", name)
    short = re.split(r"[<(]", name)[0]
    e = agg.setdefault(short, {"launches": 0, "bytes": 0.0, "us": 0.0, "variants": set()})
    e["launches"] += 1
    e["bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    e["us"] += val(r, "gpu__time_duration.sum")
    e["variants"].add(name[:90])
out = {}
for k, e in agg.items():
    out[k] = {"launches": e["launches"], "dram_bytes_per_launch": e["bytes"] / e["launches"],
              "time_us_per_launch": e["us"] / e["launches"], "dram_gbs": e["bytes"] / e["us"] / 1e3 if e["us"] else None,
              "variants": sorted(e["variants"])}
json.dump(out, sys.stdout, indent=1)
