"""Per-kernel SASS opcode counts of libsgx.so (static: instructions in the cubin, not executed counts).
usage: python tools/sass_opcodes.py [libsgx.so] > profiles/sass_opcodes.txt
What to look for (B200_PROFILING.md): UBLKCP / SYNCS = TMA bulk copy + mbarrier; FFMA2 / FADD2 / FMUL2 = the packed
FP32 pairs of sm_100; SHFL = warp shuffles; LDS / STS = shared-memory traffic; UTMALDG / UTC*MMA / LDTM would be
tensor-map TMA and tcgen05 (not used: see DESIGN.md section 3)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "multi-spectrogram-viewer_b200/libsgx.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "UBLKCP", "SYNCS", "ATOMS",
        "HMMA", "UTMALDG", "UTMASTG", "LDTM"]
rows, cur, arch = {}, None, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"sgx::\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(sgx::[A-Za-z]+(Launch)?( const)?(, [^)]*)?\)", "", cur).replace("void ", "")
        rows[cur] = collections.Counter()
        continue
    m = re.search(r"\.target\s+(sm_\w+)|arch = (sm_\w+)", line)
    if m:
        arch = m.group(1) or m.group(2)
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        rows[cur][m.group(1)] += 1
        rows[cur]["_total"] += 1
print(f"# cuobjdump -sass {lib}  (arch {arch}); static instruction counts per kernel")
print(f"{'kernel':78s} {'total':>7s} " + " ".join(f"{c:>7s}" for c in cols))
for k in sorted(rows, key=lambda k: -rows[k]["_total"]):
    c = rows[k]
    utc = sum(v for o, v in c.items() if o.startswith("UTC") and "MMA" in o)
    print(f"{k[:78]:78s} {c['_total']:7d} " + " ".join(f"{c[x]:7d}" for x in cols) + (f"  UTC*MMA={utc}" if utc else ""))
tot = collections.Counter()
for c in rows.values():
    tot.update(c)
print(f"{'ALL KERNELS':78s} {tot['_total']:7d} " + " ".join(f"{tot[x]:7d}" for x in cols))
