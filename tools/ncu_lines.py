#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel: joins the SASS page of an .ncu-rep
(ncu -i REP --page source --csv) with the line table of the cubin inside the library (nvdisasm -g).

    python tools/ncu_lines.py REP.ncu-rep KERNEL_SUBSTRING [LIB.so]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, pat = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ocpg_b200", "lib", "libmsda_sm100.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]
iA, iS, iI, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(r[iS].strip(), int(r[iI] or 0), int(r[iSm] or 0)) for r in rows[2:] if len(r) > iI]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
# find the function whose demangled name matches the profiled kernel
mangled = None
norm = lambda t: re.sub(r"\s+", "", t.replace("(int)", "").replace("(bool)", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
                        .replace("void ", "").replace("false", "0").replace("true", "1"))
head = lambda t: re.match(r"^(\w+(<.*?>)?)\(", norm(t)).group(1) if re.match(r"^(\w+(<.*?>)?)\(", norm(t)) else norm(t)
want = head(kname)
for m in re.finditer(r"\.section\s+\.text\.(\S+?),", dis):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    if pat in name and head(name) == want:
        mangled = m.group(1); break
if mangled is None:
    sys.exit("kernel not found in the cubin: " + kname)
body = dis.split(".text." + mangled + ":")[1].split(".section")[0]
line = None; per_instr = []
for ln in body.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m:
        per_instr.append((line, m.group(2).strip()))
if len(per_instr) != len(sass):
    print(f"warning: {len(per_instr)} instructions in the cubin vs {len(sass)} in the report", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
for (ln, _), (_, n, smp) in zip(per_instr, sass):
    a = agg[ln]; a[0] += n; a[1] += smp; a[2] += 1
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"{kname[:100]}\n total warp instructions {tot_i}, samples {tot_s}")
src_cache = {}
def text(ln):
    if ln is None: return ""
    f, n = ln
    for d in ("ocpg_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            src_cache.setdefault(p, open(p).read().splitlines())
            L = src_cache[p]
            return L[n - 1].strip()[:90] if n <= len(L) else ""
    return ""
for ln, (n, smp, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", 45))]:
    print(f"{100*n/tot_i:5.1f}% inst {100*smp/max(tot_s,1):5.1f}% smp {k:4d} sass  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<4d} {text(ln)}")
