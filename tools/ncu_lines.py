"""Aggregate an ncu SASS source page per CUDA source line.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <object.o> <mangled-name substring>"""
import csv, re, subprocess, sys, tempfile, os, collections
rep, kre, obj, fn = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 35
by_samples = len(sys.argv) > 6 and sys.argv[6] == "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel block only
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]; idx = {n: i for i, n in enumerate(h)}
end = hi[1] - 1 if len(hi) > 1 else len(rows)
inst = {}
for r in rows[hi[0] + 1:end]:
    try:
        a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        inst[a] = (int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]]),
                   float(r[idx["Avg. Threads Executed"]] or 0), r[1])
    except Exception:
        pass
base = min(inst)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
lines = dis.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and fn in l][0]
cur = None; amap = {}
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m: amap[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0.0])
for a, (n, s, thr, txt) in inst.items():
    k = amap.get(a - base, ("?", 0))
    agg[k][0] += n; agg[k][1] += s; agg[k][2] += n * thr
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-inst {tot}  samples {ts}")
src = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_samples else 0])[:top]:
    if f not in src:
        try: src[f] = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)).read().splitlines()
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print(f"{v[0]:>11d} {100*v[0]/tot:5.1f}% smp {100*v[1]/max(ts,1):5.1f}% thr {v[2]/max(v[0],1):5.1f}  {f}:{ln}  {text}")
