"""aggregate an ncu SASS source page by CUDA source line using nvdisasm line info of the matching cubin
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <function substring> [source file]"""
import csv, re, subprocess, sys
from collections import defaultdict
rep, kre, cubin, fn = sys.argv[1:5]
srcfile = sys.argv[5] if len(sys.argv) > 5 else None
sel = ["--kernel-id", kre] if kre.startswith(":") else ["-k", "regex:" + kre]
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = end = None
for i, l in enumerate(dis):
    if l.startswith(".text.") and fn in l and start is None: start = i; continue
    if start is not None and l.startswith(".text."): end = i; break
end = end or len(dis)
seq = []; cur = None
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((cur, m.group(2).strip()))
rows = list(csv.reader(sass.split("\n")))
hdr = rows[1]; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    try: data.append((int(r[si]), int(r[ie]), r[1].strip()))
    except Exception: pass
print("disasm", len(seq), "ncu", len(data))
agg = defaultdict(lambda: [0, 0])
for k in range(min(len(seq), len(data))):
    agg[seq[k][0]][0] += data[k][0]; agg[seq[k][0]][1] += data[k][1]
te = sum(v[1] for v in agg.values()); ts = sum(v[0] for v in agg.values())
src = open(srcfile).read().split("\n") if srcfile else []
for ln, (s, e) in sorted(agg.items(), key=lambda kv: -(kv[1][1] / max(te, 1) + kv[1][0] / max(ts, 1)))[:30]:
    f, l = ln if ln else ("?", 0)
    text = src[l - 1].strip()[:100] if src and srcfile.endswith(f) and l > 0 else ""
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * e / max(te, 1), 100 * s / max(ts, 1), f, l, text))
