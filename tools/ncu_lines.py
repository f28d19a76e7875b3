"""Join an ncu SASS source-page CSV with nvdisasm line info (dev tool).
usage: python tools/ncu_lines.py <src.csv from `ncu -i rep --page source --csv`> <kernel substring> [top]"""
import csv, re, subprocess, sys, os, tempfile, collections
src_csv, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rrt_mpc_b200", "libcudampc.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# collect (offset -> (file,line)) for the kernel
loc = {}; cur = None; infn = False
for ln in dis:
    if ln.startswith(".text."):
        infn = kern in ln
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16) if data[0][ix["Address"]].startswith("0x") else int(data[0][ix["Address"]])
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
ti = ts = 0
for r in data:
    a = r[ix["Address"]]; a = int(a, 16) if a.startswith("0x") else int(a)
    off = a - base
    l = loc.get(off, (None, "?"))[0]
    n = float(r[ix["Instructions Executed"]] or 0); s = float(r[ix["# Samples"]] or 0)
    agg[l][0] += n; agg[l][1] += s; ti += n; ts += s
    op = r[ix["Source"]].split()
    op = (op[1] if op and op[0].startswith("@") else (op[0] if op else "?")).split(".")[0]
    agg[l][2][op] += n
print(f"total inst {ti:.3g} samples {ts:.0f}")
for l, (n, s, ops) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{str(l):40s} inst {100*n/ti:5.1f}%  samples {100*s/ts:5.1f}%   {dict(ops.most_common(4))}")
# per-function summary: map (file, line) -> enclosing function via a crude scan of the sources
import bisect
csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rrt_mpc_b200", "csrc")
fmap = {}
for fn in os.listdir(csrc):
    starts = []
    for i, ln in enumerate(open(os.path.join(csrc, fn)), 1):
        m = re.match(r"(?:template\s*<[^>]*>\s*)?(?:MPC_HD|__device__|__global__|static|inline)[^;(]*?([A-Za-z_0-9]+)\s*\(", ln)
        if m: starts.append((i, m.group(1)))
    fmap[fn] = starts
reg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
F64 = {"DFMA", "DMUL", "DADD", "DSETP", "DMNMX"}
for l, (n, s, ops) in agg.items():
    name = "?"
    if l and l[0] in fmap and fmap[l[0]]:
        st = fmap[l[0]]; j = bisect.bisect_right([a for a, _ in st], l[1]) - 1
        name = f"{l[0]}:{st[j][1]}" if j >= 0 else l[0]
    elif l: name = l[0]
    reg[name][0] += n; reg[name][1] += s; reg[name][2] += n + sum(v for k, v in ops.items() if k in F64)
tot_issue = sum(v[2] for v in reg.values())
print("\nper function: inst%, samples%, issue-slot% (fp64 counted twice)")
for name, (n, s, iss) in sorted(reg.items(), key=lambda kv: -kv[1][2])[:30]:
    print(f"{name:45s} inst {100*n/ti:5.1f}%  samples {100*s/ts:5.1f}%  issue {100*iss/tot_issue:5.1f}%")
