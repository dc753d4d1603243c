"""Developer tool: attribute an `ncu --page source --csv` SASS table (trimmed: Address, Source, # Samples, Instructions
Executed, stall_*) to CUDA source lines by matching instruction order against `nvdisasm -g -c` of the same cubin.

usage: python tools/sass_hotspots.py <ncu_sass.csv> <dir with *.sass from nvdisasm -g -c> [kernel substring] [top N]
"""
import csv, re, sys, collections

csv.field_size_limit(1 << 30)
path, sassdir = sys.argv[1], sys.argv[2]
filt = sys.argv[3] if len(sys.argv) > 3 else ""
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25


def load_disasm(sassdir):
    import glob
    funcs = {}
    for f in glob.glob(sassdir + "/*.sass"):
        cur, line, inl = None, None, None
        for ln in open(f, errors="replace"):
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur = m.group(1); funcs[cur] = []; continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                line = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip()); continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur is not None:
                funcs[cur].append((m.group(2).strip(), line))
    return funcs


funcs = load_disasm(sassdir)
rows = csv.reader(open(path))
kern, hdr, table = None, None, []
tables = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if kern: tables.append((kern, hdr, table))
        kern, hdr, table = r[1], None, []; continue
    if hdr is None: hdr = r; continue
    table.append(r)
if kern: tables.append((kern, hdr, table))

for kern, hdr, table in tables:
    if filt not in kern: continue
    base = re.sub(r"\(.*", "", kern.replace("void ", "")).split("<")[0].split("::")[-1]
    cands = [k for k, v in funcs.items() if base in k and len(v) == len(table)]
    print(f"===== {kern[:100]}  ({len(table)} SASS instrs, disasm candidates: {len(cands)})")
    if not cands: continue
    dis = funcs[cands[0]]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = tot_i = 0
    for r, (ins, line) in zip(table, dis):
        s, n = int(r[si] or 0), int(r[ii] or 0)
        key = (line[0], line[1]) if line else ("?", 0)
        a = agg[key]; a[0] += s; a[1] += n
        for c in stall_cols:
            v = int(r[c] or 0)
            if v: a[2][hdr[c][6:]] += v
        tot_s += s; tot_i += n
    src_cache = {}
    for (f, l), (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        if f not in src_cache:
            try: src_cache[f] = open("botorch_b200/csrc/" + f).read().split("\n")
            except OSError: src_cache[f] = []
        text = src_cache[f][l - 1].strip()[:90] if 0 < l <= len(src_cache[f]) else ""
        top = ",".join(f"{k}:{100 * v // max(s, 1)}" for k, v in st.most_common(3))
        print(f"{100 * s / max(tot_s, 1):5.1f}% smp {100 * n / max(tot_i, 1):5.1f}% ins  {f}:{l:<4} [{top}]  {text}")
