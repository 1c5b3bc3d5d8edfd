"""Turn the raw ncu outputs of scripts/gpu_final.sh (gpurun_out/launches.csv, gpurun_out/mm_tile_full_raw.csv)
into the committed text summaries under profiles/.   python scripts/summarize_ncu.py <tag>"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_s2"
out_dir = os.path.join(ROOT, "profiles")

# ---- launch list ------------------------------------------------------------------------------------
rows = []
with open(os.path.join(ROOT, "gpurun_out", "launches.csv")) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for rec in csv.DictReader(lines):
    if rec.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((rec["Kernel Name"], rec["Grid Size"], float(rec["Metric Value"].replace(",", ""))))
agg = collections.OrderedDict()
for name, grid, ns in rows:
    short = name.split("(")[0][:60]
    key = "%s grid%s" % (short, grid.replace(" ", ""))
    cnt, tot = agg.get(key, (0, 0.0))
    agg[key] = (cnt + 1, tot + ns)
total = sum(t for _, t in agg.values())
with open(os.path.join(out_dir, tag + "_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 1 --warmup 1 --no-backward --no-cpu-baseline --nsplit 1\n")
    f.write("# first 900 launches (factorisation + eager warm-up rollouts, R=32 in one batch); cold-cache, serialised: compare SHARES\n")
    f.write("%-78s %6s %12s %7s %9s\n" % ("kernel / grid", "count", "total_ns", "share", "avg_us"))
    for key, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        f.write("%-78s %6d %12d %6.1f%% %9.1f\n" % (key[:78], cnt, tot, 100.0 * tot / total, tot / cnt / 1e3))

# ---- full capture of the tile kernel -------------------------------------------------------------------
raw = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", "mm_tile_full_raw.csv"))))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
picked = {}
for h, u, v in zip(hdr, units, vals):
    base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
    for w in want:
        if h.endswith(w) and w not in picked:
            picked[w] = (v, u)
    if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "average_warps" in h:
        picked[h.split(".", 2)[-1] if h[0].isupper() else h] = (v, u)
kname = vals[hdr.index("Kernel Name")]
with open(os.path.join(out_dir, tag + "_mm_tile_ncu_full.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:mm_tile -s 1 -c 1 python scripts/prof_mm.py 32\n")
    f.write("# %s grid %s: dynamics GP N=300 D=12 E=10, R=32 restarts\n" % (kname[:60], vals[hdr.index("Grid Size")]))
    for k in sorted(picked):
        f.write("%s = %s %s\n" % (k, picked[k][0], picked[k][1]))
to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = float(picked["dram__bytes_read.sum"][0].replace(",", "")) * to_bytes[picked["dram__bytes_read.sum"][1]]
wr = float(picked["dram__bytes_write.sum"][0].replace(",", "")) * to_bytes[picked["dram__bytes_write.sum"][1]]
json.dump({"kernel": "mm_tile_kernel<3,3>", "config": "N=300 D=12 E=10 R=32", "dram_bytes_per_launch": rd + wr,
           "source": "profiles/%s_mm_tile_ncu_full.txt" % tag}, open(os.path.join(out_dir, "r01_tile_traffic.json"), "w"))
print(open(os.path.join(out_dir, tag + "_launches.txt")).read())
print(open(os.path.join(out_dir, tag + "_mm_tile_ncu_full.txt")).read())
