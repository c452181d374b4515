"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/profile_summary.py gpurun_out/r1_full.ncu-rep gpurun_out/r1_launches.csv r1

writes profiles/<tag>_kernels.csv (one row per profiled launch, the metrics B200_PROFILING.md names),
profiles/<tag>_launches.md (per-kernel share of the step from the gpu__time_duration launch list) and
profiles/roofline_traffic.json (dram bytes per launch of the pooling kernel, read by bench.py)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def short(name):
    return name.replace("void ", "").replace("wsovod::", "").split("(")[0]


def launches_md(launches, tag, out):
    """ncu launch list (gpu__time_duration per launch) -> profiles/<tag>_launches.md: per-kernel share of the run"""
    rows = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(short(r[ki])[:70], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi]) / 1e3
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write(f"source: `{os.path.basename(launches)}`; {sum(v[0] for v in agg.values())} launches, {tot:.0f} us in total "
                "(cold-cache, serialised: compare shares, not absolutes)\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {v[1] / tot:.3f} |\n")


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [m for m in METRICS if m in idx]
    traffic = {}
    with open(os.path.join(out, f"{tag}_kernels.csv"), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["kernel"] + [f"{m} [{units[idx[m]]}]" for m in cols])
        for r in rows[2:]:
            k = short(r[idx["Kernel Name"]])
            wr.writerow([k] + [r[idx[m]] for m in cols])
            def val(m):
                v, u = float(r[idx[m]]), units[idx[m]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            tot = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            if k.startswith("roi_pool7_pyr_kernel<4>") or k.startswith("roi_pool7_pyr_kernel<(int)4>"):
                traffic["roi_pool"] = tot          # values-only pooling of bench.py: the block-max kernel
            if k.startswith("roi_pool7_kernel<4, 0>") or k.startswith("roi_pool7_kernel<(int)4, (bool)0>"):
                traffic.setdefault("roi_pool_scan", tot)
            if k.startswith("roi_pool7_kernel<4, 1>") or k.startswith("roi_pool7_kernel<(int)4, (bool)1>"):
                traffic["roi_pool+argmax"] = tot
    traffic["source"] = f"ncu --set full, {os.path.basename(rep)}, config c2, dram__bytes_read.sum + dram__bytes_write.sum per launch"
    json.dump(traffic, open(os.path.join(out, "roofline_traffic.json"), "w"), indent=1)

    launches_md(launches, tag, out)
    print("wrote", out)


if __name__ == "__main__":
    main()
