"""Turn ncu reports brought back in gpurun_out/ into the small text summaries kept under profiles/
(test infrastructure; runs in the authoring container, no GPU needed).
    python tests/summarize_profiles.py <launches.csv> <full.ncu-rep>[,<more.ncu-rep>...] <tag>"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(ROOT, "profiles")

# ---- launch list: per-kernel share of GPU time (cold-cache, serialised: compare shares, not absolutes)
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ik])
    name = name.replace("void b200::", "").replace("void ", "")
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
with open(os.path.join(out_dir, f"{tag}_launch_shares.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3\n")
    f.write(f"# {sum(cnt.values())} launches, total {total/1e6:.2f} ms (cold-cache, serialised)\n")
    f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
    for k, v in tot.most_common():
        f.write(f"{k[:70]:70s} {cnt[k]:8d} {v/1e3:12.1f} {v/1e3/cnt[k]:10.1f} {100*v/total:6.1f}%\n")

# ---- full capture: key metrics per kernel instance
reps = rep.split(",")
rep = reps[0]
rr = None
for one in reps:                      # later reports append their kernel rows (same metric set)
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if rr is None:
        rr = part
    else:
        assert part[0] == rr[0], "metric sets differ between reports"
        rr += part[2:]
h = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
idx = [h.index(w) for w in want if w in h]
seen = collections.OrderedDict()
for r in rr[2:]:
    d = {h[i]: r[i] for i in idx}
    key = re.sub(r"\(CUtensor.*", "", d["Kernel Name"])
    if key.startswith("pack_") or "copy_f32" in key or "cvt16" in key:
        continue                      # weight packing (load time), not part of the step
    seen.setdefault(key, d)
with open(os.path.join(out_dir, f"{tag}_resblock_kernels_ncu.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on : python tests/prof_step.py --B 16 --T 861 --iters 2\n")
    f.write("# first instance of each kernel specialisation; units: us, %, inst, MB\n")
    for k, d in seen.items():
        f.write(k + "\n")
        for w in want[1:]:
            if w in d:
                f.write(f"    {w:75s} {d[w]}\n")
# dominant kernel traffic (bench.py reads it)
for k, d in seen.items():
    if "resblock3_kernel<128" in k:
        def mb(x):
            return float(x.replace(",", ""))
        unit_r = rr[1][h.index("dram__bytes_read.sum")]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit_r]
        unit_w = rr[1][h.index("dram__bytes_write.sum")]
        scale_w = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit_w]
        t = mb(d["dram__bytes_read.sum"]) * scale + mb(d["dram__bytes_write.sum"]) * scale_w
        json.dump({"kernel": k, "dram_bytes_per_launch": t, "source": os.path.basename(rep),
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum, one launch at B=16,T=861"},
                  open(os.path.join(out_dir, "dominant_kernel_traffic.json"), "w"), indent=1)
        break
print(open(os.path.join(out_dir, f"{tag}_launch_shares.txt")).read())
print(open(os.path.join(out_dir, f"{tag}_resblock_kernels_ncu.txt")).read())
