"""Turns gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: make_profile_summary.py <tag> <launches.csv> <bench.json> <name=report.ncu-rep> [...]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, bench = sys.argv[1], sys.argv[2], sys.argv[3]
reports = [a.split("=", 1) for a in sys.argv[4:]]
out = [f"# ncu summary {tag}\n"]
b = json.load(open(bench))
out.append("## bench line (python bench.py, not under a profiler)\n```json\n" + json.dumps(b, indent=1) + "\n```\n")
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg[row["Kernel Name"].split("(")[0][-60:]].append(v)
tot = sum(sum(v) for v in agg.values())
out.append(f"## launch list ({os.path.basename(launches)}): `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 2 --warmup 1`\n")
out.append("Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
out.append("| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"| `{k}` | {len(v)} | {sum(v)/len(v):.2f} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% |")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__waves_per_multiprocessor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
for name, rep in reports:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out.append(f"\n## `ncu --set full` : {name} ({os.path.basename(rep)})\n")
    out.append("| metric | unit | value (per captured launch) |\n|---|---|---|")
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                out.append(f"| {w} | {units[i]} | {', '.join(r[i] for r in rows[2:])} |")
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = f"/tmp/_sass_{name}.csv"
    open(tmp, "w").write(sass)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_summary.py"), tmp, "12"], capture_output=True, text=True).stdout
    out.append("\nStall mix and hottest SASS (warp-sampling):\n```\n" + summ + "```")
path = os.path.join(ROOT, "profiles", f"{tag}_summary.md")
open(path, "w").write("\n".join(out) + "\n")
print("wrote", path)
