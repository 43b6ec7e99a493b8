"""Aggregate an ncu SASS-page CSV by CUDA source line (nvdisasm -g line info).
usage: python scripts/ncu_by_line.py <sass.csv> <nvdisasm -g -c output> <source.cu> [topN]"""
import csv, re, sys, collections
sass_csv, dis, src = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: lines like  //## File "...", line 123   followed by  /*0040*/ INSTR
addr2line, cur = {}, None
for ln in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/', ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
stall_tot = collections.Counter()
line_stall = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) <= isamp: continue
    a = int(r[ia], 16) if r[ia].startswith("0x") or re.match(r'^[0-9a-f]+$', r[ia]) else int(r[ia])
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    n = int(r[ii] or 0); s = int(r[isamp] or 0)
    agg[key][0] += n; agg[key][1] += s; agg[key][2] += 1
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: stall_tot[hdr[i]] += v; line_stall[key][hdr[i]] += v
srcl = open(src).read().split("\n")
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print("total inst", tot_i, "samples", tot_s)
print("stalls:", ", ".join(f"{k[6:]}={v}" for k, v in stall_tot.most_common(10)))
print("--- by samples")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    f, l = key
    text = srcl[l - 1].strip()[:90] if f == src.split('/')[-1] and 0 < l <= len(srcl) else f
    top = ",".join(f"{k[6:]}:{c}" for k, c in line_stall[key].most_common(3))
    print(f"{f}:{l:4d} inst={v[0]:9d} ({100*v[0]/tot_i:4.1f}%) samp={v[1]:6d} ({100*v[1]/tot_s:4.1f}%) sass={v[2]:4d} [{top}] | {text}")
if len(sys.argv) > 5:
    print("--- by line range")
    bounds = [int(x) for x in sys.argv[5].split(",")]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        ii = sum(v[0] for k, v in agg.items() if k[0] == src.split('/')[-1] and lo <= k[1] < hi)
        ss = sum(v[1] for k, v in agg.items() if k[0] == src.split('/')[-1] and lo <= k[1] < hi)
        print(f"lines {lo}-{hi}: inst={ii} ({100*ii/tot_i:.1f}%) samples={ss} ({100*ss/tot_s:.1f}%)")
    ii = sum(v[0] for k, v in agg.items() if k[0] != src.split('/')[-1]); ss = sum(v[1] for k, v in agg.items() if k[0] != src.split('/')[-1])
    print(f"other files: inst={ii} ({100*ii/tot_i:.1f}%) samples={ss} ({100*ss/tot_s:.1f}%)")
