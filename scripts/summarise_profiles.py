"""Turn gpurun_out/<round>_{launches_c5.csv, top_kernels.ncu-rep} into the tracked summaries under profiles/:
   <round>_launches_c5.csv (copied), <round>_ncu_summary.md, <round>_traffic.json (dram bytes per launch of each kernel,
   read by bench.py for roofline.traffic).  Runs in the build container (ncu -i needs no GPU)."""
import csv, json, os, shutil, subprocess, sys, collections

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PROF, exist_ok=True)


def short(name):
    name = name.replace("void ", "").replace("rr::", "").replace("tc::", "").replace("pipe::", "")
    return name.split("(")[0]


# ---- launch list -----------------------------------------------------------------------------------------------------
src = os.path.join(OUT, f"{R}_launches_c5.csv")
lines = [l for l in open(src) if not l.startswith("==")]
open(os.path.join(PROF, f"{R}_launches_c5.csv"), "w").writelines(lines)
rows = list(csv.reader(lines))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
per = collections.OrderedDict()
tot = 0.0
for r in rows[1:]:
    if len(r) <= vi:
        continue
    try:
        ns = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = short(r[ki])
    per.setdefault(k, [0, 0.0])
    per[k][0] += 1
    per[k][1] += ns
    tot += ns
md = [f"# {R} ncu summaries (B200, bench.py c5 workload, one GPU)", "",
      f"Launch list: `{R}_launches_c5.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none`, {sum(v[0] for v in per.values())} launches = warm-up + 2 steps "
      "of `bench.py --steps 2 --warmup 1`).  Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's own CUDA-event shares.", "",
      "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, (n, ns) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    if ns / tot >= 0.002:
        md.append(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f} % |")

# ---- full captures ----------------------------------------------------------------------------------------------------
import glob
traffic = {}
reps = sorted(glob.glob(os.path.join(OUT, f"{R}_full_*.ncu-rep")))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def unit_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def us(v, u):
    return float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)


groups = collections.OrderedDict()          # kernel -> list of {metric: (value, unit)}: every report is read with ITS OWN header
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    if len(part) < 3:
        continue
    hh, uu = part[0], part[1]
    kn = hh.index("Kernel Name")
    idx = {w: hh.index(w) for w in want if w in hh}
    for r in part[2:]:
        if len(r) <= kn:
            continue
        groups.setdefault(short(r[kn]), []).append({w: (r[i], uu[i]) for w, i in idx.items()})
if groups:
    md += ["", f"## `--set full` captures (`gpurun_out/{R}_full_*.ncu-rep`, launches of one steady-state step of bench.py; the longest launch of each kernel is shown)", ""]
    for k, lst in groups.items():
        durs = [us(*m["gpu__time_duration.sum"]) for m in lst]
        big = lst[max(range(len(lst)), key=lambda i: durs[i])]
        rd = sum(unit_bytes(*m["dram__bytes_read.sum"]) for m in lst) / len(lst)
        wr = sum(unit_bytes(*m["dram__bytes_write.sum"]) for m in lst) / len(lst)
        traffic[k] = {"launches_captured": len(lst), "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                      "mean_us": sum(durs) / len(durs)}
        md += [f"### `{k}`  ({len(lst)} launches captured; mean {sum(durs) / len(durs):.1f} us, DRAM {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written per launch)", "",
               "| metric | value |", "|---|---|"]
        for w in want:
            if w in big:
                md.append(f"| {w} | {big[w][0]} {big[w][1]} |")
        md.append("")
json.dump(traffic, open(os.path.join(PROF, f"{R}_traffic.json"), "w"), indent=1)
open(os.path.join(PROF, f"{R}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md[:40]))
