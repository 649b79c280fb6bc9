"""Turns the files a `gpurun` profiling call leaves in gpurun_out/ into the committed summaries under profiles/:
  gpurun_out/bench_final.json            -> profiles/<tag>_bench.json
  gpurun_out/launches_<tag>.csv          -> profiles/<tag>_launches.csv + <tag>_launches_summary.md
  gpurun_out/attn_small_<tag>.ncu-rep    -> profiles/<tag>_attn_small_kernel_{summary.json, ncu_raw.csv, ncu_details.txt}
Usage: python tools/summarise_profiles.py r1_final   (needs `ncu` on PATH to read the .ncu-rep)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_final"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

bench = os.path.join(G, "bench_final.json")
ms_step = None
if os.path.exists(bench):
    shutil.copy(bench, os.path.join(P, f"{tag}_bench.json"))
    ms_step = json.load(open(bench))["ms_per_step"]

rep = os.path.join(G, f"attn_small_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_attn_small_kernel_ncu_raw.csv"), "w").write(raw)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_attn_small_kernel_ncu_details.txt"), "w").write(det)
    r = list(csv.reader(raw.splitlines()))
    d = {h: (v, u) for h, u, v in zip(r[0], r[1], r[2])}
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors_srcunit_tex_op_read.sum"]
    out = {k: {"value": d[k][0], "unit": d[k][1]} for k in keys if k in d}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tob = lambda k: float(d[k][0].replace(",", "")) * scale[d[k][1]]
    out["dram_bytes_per_launch"] = tob("dram__bytes_read.sum") + tob("dram__bytes_write.sum")
    out["kernel"] = d["Kernel Name"][0]
    out["command"] = ("ncu --set full --clock-control none --import-source on -k regex:attn_small_kernel -s 15 -c 1 "
                      "python bench.py --steps 1 --warmup 3 --no-cpu")
    json.dump(out, open(os.path.join(P, f"{tag}_attn_small_kernel_summary.json"), "w"), indent=1)
    print({k: out[k]["value"] for k in keys[:8] if k in out})

lst = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lst):
    shutil.copy(lst, os.path.join(P, f"{tag}_launches.csv"))
    lines = [l for l in open(lst) if not l.startswith("==")]
    tot, seq = collections.defaultdict(lambda: [0, 0.0]), []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        for junk in ("void ", "hn::", "<unnamed>::", "unnamed>::"):
            name = name.replace(junk, "")
        tot[name][0] += 1
        tot[name][1] += v
        seq.append((name, v))
    T = sum(v[1] for v in tot.values())
    idx = [i for i, (k, _) in enumerate(seq) if "head_kernel" in k]
    a, b = idx[-2] + 1, idx[-1] + 1
    with open(os.path.join(P, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# {tag} — launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` under ncu\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "
                f"gpurun_out/launches_{tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu`\n")
        f.write(f"Raw per-launch list: `profiles/{tag}_launches.csv` ({len(seq)} launches = one weight-packing pass + 9 "
                "forwards). Per-launch times under ncu are cold-cache and serialised (programmatic dependent launch "
                "cannot overlap anything there): compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / T:.2f}% |\n")
        f.write(f"| **total** | {len(seq)} | {T:.3f} | 100% |\n\n")
        f.write(f"One forward = {b - a} launches, {sum(v for _, v in seq[a:b]):.3f} ms summed under ncu")
        if ms_step:
            f.write(f" (bench.py, same code: {ms_step:.2f} ms per step with launches overlapped)")
        f.write(".\n")
    print(open(os.path.join(P, f"{tag}_launches_summary.md")).read()[-1700:])

tl = os.path.join(G, f"{tag}_train_launches.csv")
if os.path.exists(tl):
    # launch list of the training step: `ncu --metrics gpu__time_duration.sum ... python bench.py --mode train --steps 1 --warmup 1`
    shutil.copy(tl, os.path.join(P, f"{tag}_train_launches.csv"))
    lines = [l for l in open(tl) if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        for junk in ("void ", "hn::", "<unnamed>::", "unnamed>::"):
            name = name.replace(junk, "")
        tot[name][0] += 1
        tot[name][1] += v
        n += 1
    T = sum(v[1] for v in tot.values())
    with open(os.path.join(P, f"{tag}_train_launches_summary.md"), "w") as f:
        f.write(f"# {tag} — launch list of the training step under ncu\n\n")
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file "
                f"gpurun_out/{tag}_train_launches.csv python bench.py --mode train --steps 1 --warmup 1 --no-cpu` "
                f"({n} launches: weight packing, the bench's warm-up and timed training steps and its forward-only "
                "passes). Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| `{k[:70]}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / T:.2f}% |\n")
        f.write(f"| **total (all kernels)** | {n} | {T:.3f} | 100% |\n")
    print("train launch summary written")
