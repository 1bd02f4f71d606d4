"""Join an ncu report's per-SASS-instruction metrics with source lines (via nvdisasm -g).

usage: python tools/ncu_lines.py <report.ncu-rep> <cubin> [min_pct]
Prints, per source line: warp instructions executed (% of kernel), stall samples (%), average
active threads, dominant stall reasons.  Used to write the summaries under profiles/.
"""
import csv, re, subprocess, sys, collections, io

rep, cubin = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
func = sys.argv[4] if len(sys.argv) > 4 else None      # substring of the (mangled) kernel name when the cubin holds several
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur = None
in_func = func is None
for ln in sass.splitlines():
    if ln.lstrip().startswith(".section") and ".text." in ln:
        in_func = func is None or func in ln
        cur = None
        continue
    if not in_func:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
hi = his[0]                                             # first captured launch of the report
hdr = rows[hi]
ix = {n: i for i, n in enumerate(hdr)}
end = his[1] if len(his) > 1 else len(rows)
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0] != "Address"]
base = int(data[0][0], 16)
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: dict(inst=0, thr=0, samp=0, st=collections.Counter()))
tot_i = tot_s = 0
for r in data:
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    inst = int(r[ix["Instructions Executed"]] or 0)
    a["inst"] += inst
    a["thr"] += int(r[ix["Thread Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    a["samp"] += s
    tot_i += inst
    tot_s += s
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            a["st"][c[6:]] += v
print(f"total warp-instructions {tot_i}  samples {tot_s}")
src_cache = {}
def src(f, l):
    try:
        if f not in src_cache:
            import glob
            p = glob.glob(f"/root/repo/fast_limo_b200/csrc/{f}")
            src_cache[f] = open(p[0]).read().splitlines() if p else []
        return src_cache[f][l - 1].strip()[:70]
    except Exception:
        return ""
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"]):
    pi, ps = 100 * a["inst"] / max(tot_i, 1), 100 * a["samp"] / max(tot_s, 1)
    if pi < min_pct and ps < min_pct:
        continue
    st = ",".join(f"{k}:{100*v/max(a['samp'],1):.0f}" for k, v in a["st"].most_common(3))
    print(f"{key[0]}:{key[1]:<4d} inst {pi:5.1f}%  samp {ps:5.1f}%  thr/warp {a['thr']/max(a['inst'],1):5.1f}  [{st}]  {src(*key)}")
