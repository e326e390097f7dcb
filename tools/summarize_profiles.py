"""Turns the ncu artefacts under gpurun_out/ into the committed summaries under profiles/ (run in the build container, no GPU):
    python tools/summarize_profiles.py <launches.csv> <full.ncu-rep> <tag>"""
import collections
import csv
import json
import subprocess
import sys

launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v = v / 1000 if r[iu] == "ns" else v * 1000 if r[iu] == "ms" else v
    agg[r[ik].split("(")[0].replace("void ", "")].append(v)
tot = sum(sum(v) for v in agg.values())
with open(f"profiles/{tag}_ncu_launches.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 ... python bench.py --steps 1 --warmup 1 --diffusion-steps 12 "
            "--no-cpu-baseline --no-other-configs --profile-stride 0   (tools/gpu_r02_base.sh; cold-cache, serialised launches: compare SHARES, not absolutes)\n")
    f.write(f"{'kernel':60s} {'launches':>8s} {'avg_us':>9s} {'total_us':>10s} {'share':>7s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k[:60]:60s} {len(v):8d} {sum(v)/len(v):9.2f} {sum(v):10.1f} {sum(v)/tot*100:6.1f}%\n")
print(open(f"profiles/{tag}_ncu_launches.txt").read())

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
keep = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed.sum"]
out = {}
with open(f"profiles/{tag}_ncu_full_summary.txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:'encoder_stack' -s 2 -c 1 python tools/profile_layer.py\n"
            "(cfg2: B=256, L=256; one score evaluation = one launch of the persistent encoder-stack kernel; under ncu, so durations are NOT bench values)\n\n")
    for vals in rr[2:]:
        d = dict(zip(h, vals))
        name = d.get("Kernel Name", "?").split("(")[0]
        f.write(f"== {name}\n")
        rec = {}
        for k in keep:
            if k in d:
                f.write(f"   {k:85s} {d[k]:>16s} {units[h.index(k)]}\n")
                rec[k] = (d[k], units[h.index(k)])
        out[name] = rec
        f.write("\n")
print(open(f"profiles/{tag}_ncu_full_summary.txt").read())
def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
traffic = {}
for name, rec in out.items():
    if "dram__bytes_read.sum" in rec:
        traffic[name] = to_bytes(*rec["dram__bytes_read.sum"]) + to_bytes(*rec["dram__bytes_write.sum"])
json.dump({"dram_bytes_per_launch": traffic, "source": f"profiles/{tag}_ncu_full_summary.txt"}, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
