"""Static SASS evidence per kernel of liba2v_sm100.so: counts of the Blackwell-specific mnemonics (UTCHMMA/UTCQMMA =
tcgen05.mma, UTMALDG/UTMASTG = TMA, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UBLKCP = bulk
copy, REDG/RED/ATOMG = atomics), registers and spill bytes. Run where the library is built (no GPU needed)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "animal2vec_b200", "liba2v_sm100.so")
MN = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "SYNCS", "MUFU", "REDG", "ATOMG", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
kern = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        kern[cur]["_total"] += 1
        op = m.group(1)
        for k in MN:
            if op.startswith(k):
                kern[cur][k] += 1
def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("| kernel | SASS instr | regs | local B | " + " | ".join(MN) + " |")
print("|---|---|---|---|" + "---|" * len(MN))
for k, c in kern.items():
    d = re.sub(r"\(.*", "", demangle(k)).replace("a2v::", "").replace("void ", "").replace("__nv_bfloat16", "bf16")
    r = regs.get(k, (0, 0, 0))
    print(f"| `{d}` | {c['_total']} | {r[0]} | {r[2]} | " + " | ".join(str(c[m]) if c[m] else "" for m in MN) + " |")
