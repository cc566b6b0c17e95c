"""Per-source-line profile of one kernel of an .ncu-rep: joins the SASS page of the report (instructions executed, stall
samples) with the line table of the cubin (nvdisasm --print-line-info).
usage: python tools/ncu_lines.py REP OBJ MANGLED_KERNEL_SUBSTRING [kernel-occurrence-in-report]"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, obj, sub = sys.argv[1:4]
occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# line table of the kernel: offset -> (line, inlined-from chain ignored)
lines = {}
inside = False
cur = None
for l in dis:
    if l.startswith(".text."):
        inside = sub in l
        name = l
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
demangled_key = sys.argv[5] if len(sys.argv) > 5 else None
sel = [i for i in starts if (demangled_key is None or demangled_key in rows[i][1])]
st = sel[occ]
en = min([i for i in starts if i > st] + [len(rows)])
hdr = rows[st + 1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(h, k) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[st + 2][ia], 16)
agg = collections.OrderedDict()
tot_i = tot_s = 0
for r in rows[st + 2:en]:
    off = int(r[ia], 16) - base
    key = lines.get(off, ((None, 0), ""))[0]
    a = agg.setdefault(key, [0, 0, collections.Counter()])
    a[0] += int(r[ie]); a[1] += int(r[isamp])
    for h, k in stall_cols: a[2][h] += int(r[k])
    tot_i += int(r[ie]); tot_s += int(r[isamp])
print(rows[st][1][:150])
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
src = {}
for key in sorted(agg, key=lambda k: (k[0] or "", k[1])):
    a = agg[key]
    if a[0] < tot_i * 0.002 and a[1] < tot_s * 0.002: continue
    f, ln = key
    if f and f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src.get(f, [])[ln - 1].strip()[:90] if f and ln and ln <= len(src.get(f, [])) else ""
    top = ", ".join("%s %d" % (h[6:], c) for h, c in a[2].most_common(3) if c)
    print("%-22s %5.1f%% instr %5.1f%% samples | %-90s | %s" % ("%s:%d" % (f, ln), 100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, text, top))
