"""profiles/sass_summary.txt: per kernel of libhydrotools_b200.so, how many of the SASS
instructions that matter here it holds -- TMA loads (UTMALDG), mbarrier traffic (SYNCS),
shared-memory atomics (ATOMS.*; ATOMS.CAST.SPIN = compare-and-swap loop, what fp64 / 64-bit
shared atomicAdd compiles to on sm_100a), global atomics / reductions (ATOMG, RED),
128-bit shared / global accesses.  `cuobjdump -sass` of the built library.
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "hydro-tools_b200", "hydrotools", "cuda", "_lib", "libhydrotools_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = [("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS\."), ("ATOMS.CAST.SPIN", r"ATOMS\.CAST\.SPIN"),
        ("ATOMS (native)", r"\bATOMS\.(?!CAST)"), ("ATOMG", r"\bATOMG\."), ("RED", r"\bRED\."),
        ("LDS.128", r"\bLDS\.128"), ("STG.128/LDG.128", r"\b(STG|LDG)\.E\.(\w+\.)*128"),
        ("MUFU", r"\bMUFU\."), ("DADD/DMUL/DFMA", r"\bD(ADD|MUL|FMA)\b"), ("HMMA/UTCMMA", r"\b(HMMA|UTC\w*MMA)")]
kern, rows = None, collections.OrderedDict()
arch = None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\w+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = collections.Counter(instr=0)
        continue
    if kern and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        rows[kern]["instr"] += 1
        for name, p in pats:
            if re.search(p, line):
                rows[kern][name] += 1


def short(k):
    try:
        d = subprocess.run(["cu++filt", k], capture_output=True, text=True).stdout.strip()
    except Exception:
        d = k
    d = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void ", "", d)
    if d.endswith(")"):  # drop the parameter list (the last balanced parenthesis group)
        depth = 0
        for k in range(len(d) - 1, -1, -1):
            depth += d[k] == ")"
            depth -= d[k] == "("
            if depth == 0:
                d = d[:k]
                break
    return d[:58]


names = [n for n, _ in pats]
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}  (arch {arch})")
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{n:>16s}" for n in names))
for k, c in rows.items():
    print(f"{short(k):58s} {c['instr']:6d} " + " ".join(f"{c[n]:16d}" for n in names))
tot = collections.Counter()
for c in rows.values():
    tot.update(c)
print(f"{'TOTAL':58s} {tot['instr']:6d} " + " ".join(f"{tot[n]:16d}" for n in names))
