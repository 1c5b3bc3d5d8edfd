"""Turn the raw outputs of scripts/profile_round.sh (gpurun_out/<tag>/) into the committed summaries under profiles/:
   <tag>_bench_metric.json, <tag>_bench_metric_reference.json, <tag>_launches_fwd_bwd.txt, <tag>_launches_fwd_r1.txt,
   <tag>_<kernel>_ncu_full.txt (selected raw metrics of each ncu --set full capture), <tag>_tile_traffic.json, <tag>_clocks.txt
usage: python scripts/summarize_round.py r02"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = os.path.join(ROOT, "gpurun_out", tag)
out = os.path.join(ROOT, "profiles")

for fn in ("bench_metric.json", "bench_metric_reference.json"):
    if os.path.exists(os.path.join(src, fn)):
        shutil.copy(os.path.join(src, fn), os.path.join(out, "%s_%s" % (tag, fn)))


def launches(csvname, txtname, header):
    path = os.path.join(src, csvname)
    if not os.path.exists(path):
        return
    rows = []
    for rec in csv.DictReader([ln for ln in open(path) if ln.startswith('"')]):
        if rec.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((re.sub(r"\(.*", "", rec["Kernel Name"]), rec["Grid Size"].replace(" ", ""), float(rec["Metric Value"].replace(",", ""))))
    idx = [i for i, r in enumerate(rows) if "ro_state" in r[0]]
    second = rows[idx[len(idx) // 2]:] if idx else rows              # the profiled (second) pass
    agg = collections.OrderedDict()
    for name, grid, ns in second:
        key = "%s grid%s" % (name[:58], grid)
        c, t = agg.get(key, (0, 0.0))
        agg[key] = (c + 1, t + ns)
    total = sum(t for _, t in agg.values())
    with open(os.path.join(out, txtname), "w") as f:
        f.write(header)
        f.write("%-78s %6s %12s %7s %9s\n" % ("kernel / grid", "count", "total_ns", "share", "avg_us"))
        for key, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-78s %6d %12d %6.1f%% %9.1f\n" % (key[:78], c, t, 100.0 * t / total, t / c / 1e3))
        f.write("total %.1f us\n" % (total / 1e3))


launches("launches_fwd_bwd.csv", tag + "_launches_fwd_bwd.txt",
         "# ncu --metrics gpu__time_duration.sum --clock-control none  python scripts/profile_step.py --H 3   (metric config, R=32 in ONE batch, taped forward + reverse sweep, second pass)\n"
         "# cold-cache, serialised: compare SHARES, not absolutes\n")
launches("launches_fwd_r1.csv", tag + "_launches_fwd_r1.txt",
         "# ncu --metrics gpu__time_duration.sum --clock-control none  python scripts/profile_step.py --R 1 --H 3 --no-backward   (metric shape, ONE restart, forward)\n"
         "# cold-cache, serialised\n")

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]
for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
    base = rep[:-8]
    raw = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = dict(zip(hdr, zip(units, vals)))
    with open(os.path.join(out, "%s_%s_ncu_full.txt" % (tag, base)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:%s ... python scripts/profile_step.py (metric config, R=32)\n" % base)
        f.write("# kernel: %s\n" % rec.get("Kernel Name", ("", ""))[1])
        for k in WANT:
            if k in rec:
                f.write("%-78s %-16s %s\n" % (k, rec[k][0], rec[k][1]))
        for k in sorted(rec):
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                try:
                    if float(rec[k][1]) >= 0.05:
                        f.write("%-78s %-16s %s\n" % (k, rec[k][0], rec[k][1]))
                except ValueError:
                    pass
    if base == "mm_tile":
        def by(name):
            u, v = rec[name]
            v = float(v)
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        json.dump({"kernel": rec.get("Kernel Name", ("", ""))[1], "dram_bytes_per_launch": by("dram__bytes_read.sum") + by("dram__bytes_write.sum"),
                   "source": "profiles/%s_mm_tile_ncu_full.txt" % tag}, open(os.path.join(out, tag + "_tile_traffic.json"), "w"))
if os.path.exists(os.path.join(src, "clocks.csv")):
    lines = [ln.strip() for ln in open(os.path.join(src, "clocks.csv")) if ln.strip()]
    sm = sorted(int(re.sub(r"\D", "", ln.split(",")[1])) for ln in lines[1:] if re.search(r"\d", ln.split(",")[1]))
    flags = collections.Counter()
    for ln in lines[1:]:
        parts = [p.strip() for p in ln.split(",")]
        for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
            if v.lower().startswith("active"):
                flags[nm] += 1
    with open(os.path.join(out, tag + "_clocks.txt"), "w") as f:
        busy = [v for v in sm if v >= 0.9 * sm[-1]]
        f.write("nvidia-smi -lms 200 over the WHOLE bench.py run (metric config; the run includes set-up and the CPU legs, during which the GPU idles at "
                "%d MHz): %d samples, %d of them at >= 90 %% of the max clock %d MHz; throttle reasons active in any sample: %s.  "
                "The clocks DURING the timed regions are sampled by bench.py itself (JSON key `clocks`: median under load / max / reasons).\n"
                % (sm[0], len(sm), len(busy), sm[-1], dict(flags) or "none"))
print("profiles/ updated for", tag)
