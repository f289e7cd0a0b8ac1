"""Markdown table (one row per kernel variant, averaged over its profiled launches) from `ncu --page raw --csv` exports."""
import collections, csv, re, sys
COLS = [("ms", "gpu__time_duration.sum"), ("DRAM MB", None), ("DRAM GB/s", None),
        ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("xu %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("fma %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("warps %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("smem KB", "launch__shared_mem_per_block_dynamic"),
        ("grid", "launch__grid_size"), ("block", "launch__block_size")]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "ns": 1e-6, "us": 1e-3, "ms": 1}
print("| kernel | n | " + " | ".join(c for c, _ in COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    def val(r, name):
        if name not in ix or r[ix[name]] in ("", "n/a"):
            return None
        return float(r[ix[name]].replace(",", "")) * SCALE.get(units[ix[name]].split("/")[0], 1)
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[ix["Kernel Name"]])
        name = re.sub(r"^a2v::", "", name)
        name = re.sub(r"\(.*", "", name).replace("__nv_bfloat16", "bf16")
        agg.setdefault(name, []).append(r)
    for name, rs in agg.items():
        out = []
        ms = sum(val(r, "gpu__time_duration.sum") for r in rs) / len(rs)
        by = sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in rs) / len(rs)
        for c, k in COLS:
            if c == "DRAM MB":
                out.append(f"{by / 1e6:.0f}")
            elif c == "DRAM GB/s":
                out.append(f"{by / ms / 1e6:.0f}")
            else:
                vs = [val(r, k) for r in rs]
                vs = [v for v in vs if v is not None]
                v = sum(vs) / len(vs) if vs else None
                if v is None:
                    out.append("-")
                elif c == "ms":
                    out.append(f"{v:.3f}")
                elif c == "smem KB":
                    out.append(f"{v / 1e3:.1f}")
                elif c in ("regs", "grid", "block"):
                    out.append(f"{v:.0f}")
                else:
                    out.append(f"{v:.1f}")
        print(f"| `{name}` | {len(rs)} | " + " | ".join(out) + " |")
