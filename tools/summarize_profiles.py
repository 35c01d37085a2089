"""Turn gpurun_out/ ncu artefacts into the small, committed summaries under profiles/.

    python tools/summarize_profiles.py <tag> <launches.csv> <prof.ncu-rep>

Writes profiles/<tag>_launches.csv (per-kernel aggregate of the launch list), profiles/<tag>_ncu_full_summary.json
(selected --set full counters per captured launch) and updates profiles/traffic.json (DRAM bytes per launch of the
forward kernel, read by bench.py for roofline.traffic).  Runs on the CPU box (ncu -i needs no GPU)."""
import collections, csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)


def short(name):
    name = name.replace("kd::", "").replace("fused::", "").replace("(bool)", "").replace("void ", "")
    return name.split("(CUtensorMap")[0][:90]


# ---- launch list ----
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.OrderedDict()
for x in csv.DictReader(lines):
    if x.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(x["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(x["Metric Unit"], 1.0)
    agg.setdefault(short(x["Kernel Name"]), []).append(v)
total_ours = sum(sum(v) for k, v in agg.items() if k.startswith("kd_"))
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "avg_us", "min_us", "max_us", "share_of_kd_kernels"])
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        w.writerow([k, len(v), f"{sum(v):.1f}", f"{sum(v)/len(v):.1f}", f"{min(v):.1f}", f"{max(v):.1f}",
                    f"{sum(v)/total_ours:.3f}" if k.startswith("kd_") else ""])

# ---- full capture ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = {k: hdr.index(k) for k in keys if k in hdr}
out = []
traffic = {}
for r in rows[2:]:
    d = {k: (r[i] + (" " + units[i] if units[i] else "")) for k, i in idx.items()}
    d["Kernel Name"] = short(r[idx["Kernel Name"]])
    out.append(d)

    def num(k):
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]].lower()
        return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)

    if "FwdEpi" in d["Kernel Name"] and "dram__bytes_read.sum" in idx:
        traffic["kd_umma_kernel_fwd_dram_bytes_per_launch"] = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        traffic["source"] = f"profiles/{tag}_ncu_full_summary.json (ncu --set full, one launch at BASELINE configs[1] size)"
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json"), "w"), indent=1)
if traffic:
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv")).read())
print(json.dumps(out, indent=1)[:200], "...")
print(traffic)
