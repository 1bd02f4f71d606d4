"""Turns a `ncu --set full` report of the fused kernel (+ the launch list of the same command) into the summaries committed
under profiles/: metric table, per-line hot spots, launch summary, SASS histogram + listing, roofline.traffic json.

usage: python tools/profile_summaries.py <report.ncu-rep> <launches.csv> <tag>      (tag e.g. r2c -> profiles/k1_r2c_*.txt)
"""
import collections, csv, gzip, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
so = os.path.join(ROOT, "fast_limo_b200", "libflimo_cuda.so")
out = os.path.join(ROOT, "profiles")
WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread launch__shared_mem_per_block_static
sm__cycles_active.avg sm__cycles_active.max sm__cycles_active.min sm__cycles_elapsed.max smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio smsp__issue_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active l1tex__t_sector_hit_rate.pct
lts__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed dram__bytes_read.sum
dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sectors_srcunit_tex_op_read.sum
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_membar_per_issue_active.ratio""".split()

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# ncu --set full --clock-control none --import-source on, FLIMO_PERSISTENT=0 python bench.py --steps 20 --warmup 3 --repeats 1 (c2; launches {', '.join(r[hdr.index('ID')] for r in rows[2:])} of the capture window)"]
dram = None
for r in rows[2:]:
    lines.append(f"## {r[hdr.index('Kernel Name')]}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append("%-100s %18s %s" % (w, r[i], units[i]))
    if dram is None:
        def val(name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[units[i]]
        dram = (val("dram__bytes_read.sum"), val("dram__bytes_write.sum"))
open(os.path.join(out, f"k1_{tag}_ncu_metrics.txt"), "w").write("\n".join(lines) + "\n")
json.dump({"kernel": "match_reduce_kernel<true,false>", "dram_bytes_read": int(dram[0]), "dram_bytes_write": int(dram[1]),
           "dram_bytes_per_launch": int(dram[0] + dram[1]),
           "source": f"ncu --set full capture {tag} (profiles/k1_{tag}_ncu_metrics.txt), c2 config, one launch per pass",
           "algorithmic_bytes_per_launch": 131072 * 528}, open(os.path.join(out, "k1_dram_bytes.json"), "w"), indent=1)

# per-line hot spots
tmp = "/tmp/flimo_cubins"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, "match_kernel.sm_100a.cubin")
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, cubin, "0.7", "match_reduce_kernelILb1ELb0"], capture_output=True, text=True).stdout
open(os.path.join(out, f"k1_{tag}_hotlines.txt"), "w").write(hot)

# SASS histogram + listing of the kernel
sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
blocks = re.split(r"(?=\t*Function : )", sass)
mine = next(b for b in blocks if "match_reduce_kernelILb1ELb0" in b.split("\n", 1)[0])
ops = collections.Counter(m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", mine))
with open(os.path.join(out, f"k1_{tag}_sass_histogram.txt"), "w") as f:
    f.write(f"# SASS of match_reduce_kernel<true,false> (sm_100a), capture {tag}: mnemonic histogram ({sum(ops.values())} instructions)\n")
    for k, v in ops.most_common():
        f.write("%7d %s\n" % (v, k))
with gzip.open(os.path.join(out, f"k1_{tag}_sass.txt.gz"), "wt") as f:
    f.write(mine)

# launch list
d = collections.defaultdict(list)
for r in csv.reader(open(launches)):
    if len(r) > 10 and r[0].isdigit():
        d[r[4]].append(float(r[-1]))
tot = sum(sum(v) for v in d.values())
with open(os.path.join(out, f"launches_{tag}_summary.md"), "w") as f:
    f.write(f"# Launch list, capture {tag}\n\n`FLIMO_PERSISTENT=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 python bench.py --steps 40 --warmup 3 "
            "--repeats 1 --no-cpu --no-extra`\n(one launch per pass, host filter step: the only mode ncu can replay — the resident tiles / filter kernels wait for each other).\n\n"
            "| kernel | launches | mean device time | min / max | share of device time |\n|---|---|---|---|---|\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} µs (cold-cache, serialised) | {min(v)/1e3:.2f} / {max(v)/1e3:.2f} µs | {100*sum(v)/tot:.1f} % |\n")
    v = max(d.values(), key=len)
    n3 = len(v) // 3 * 3
    per = [sum(v[i:n3:3]) / (n3 // 3) / 1e3 for i in range(3)]
    f.write(f"\nBy pass of a registration (init pose, then two near-converged poses): {per[0]:.1f} / {per[1]:.1f} / {per[2]:.1f} µs.\n")
print("written profiles/*", tag)
