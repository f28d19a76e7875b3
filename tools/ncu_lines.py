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
