"""`ncu --set full` report -> a compact per-kernel table of the sections the roofline discussion uses
(profiles/<tag>_details.csv):  python tools/profile_details.py gpurun_out/r2_pool_full.ncu-rep r2_pool"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("GPU Speed Of Light Throughput", "Memory Workload Analysis", "Warp State Statistics", "Launch Statistics",
        "Occupancy", "Scheduler Statistics", "Compute Workload Analysis")


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    ix = {n: i for i, n in enumerate(h)}
    out = os.path.join(ROOT, "profiles", f"{tag}_details.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "section", "metric", "unit", "value"])
        for r in rows[1:]:
            if len(r) <= ix["Metric Value"] or r[ix["Section Name"]] not in KEEP or not r[ix["Metric Name"]]:
                continue
            k = r[ix["Kernel Name"]].replace("void ", "").replace("wsovod::", "").split("(")[0]
            w.writerow([r[ix["ID"]], k, r[ix["Grid Size"]], r[ix["Block Size"]], r[ix["Section Name"]], r[ix["Metric Name"]],
                        r[ix["Metric Unit"]], r[ix["Metric Value"]]])
    print("wrote", out)


if __name__ == "__main__":
    main()
