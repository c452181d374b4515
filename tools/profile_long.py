"""Long-format ncu CSV (`ncu --metrics ... --csv --log-file x.csv`, one row per launch and metric) -> the committed
wide table profiles/<tag>_kernels.csv (one row per launch) and profiles/roofline_traffic.json.

    python tools/profile_long.py gpurun_out/r2_kernels_long.csv r2"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    return name.replace("void ", "").replace("wsovod::", "").split("(")[0]


def main():
    src, tag = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    c = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    units = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        key = (int(r[c["ID"]]), short(r[c["Kernel Name"]]))
        launches.setdefault(key, {})[r[c["Metric Name"]]] = r[c["Metric Value"]].replace(",", "")
        units[r[c["Metric Name"]]] = r[c["Metric Unit"]]
    metrics = list(units)
    out = os.path.join(ROOT, "profiles", f"{tag}_kernels.csv")
    traffic = {}
    with open(out, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["id", "kernel"] + [f"{m} [{units[m]}]" for m in metrics])
        for (i, k), m in launches.items():
            wr.writerow([i, k] + [m.get(x, "") for x in metrics])

            def val(x):
                v, u = float(m.get(x, 0) or 0), units.get(x, "").lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            tot = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            if k.startswith("roi_pool7_pyr_kernel<4, 0>") or k.startswith("roi_pool7_pyr_kernel<(int)4, (bool)0>"):
                traffic.setdefault("roi_pool", tot)             # values-only pooling of bench.py: block-max planes
            if k.startswith("roi_pool7_pyr_kernel<2, 1>") or k.startswith("roi_pool7_pyr_kernel<(int)2, (bool)1>"):
                traffic.setdefault("roi_pool+argmax", tot)      # block-max planes of (value, index) pairs
            if k.startswith("roi_pool7_kernel<4, 0>") or k.startswith("roi_pool7_kernel<(int)4, (bool)0>"):
                traffic.setdefault("roi_pool_scan", tot)
            if k.startswith("roi_pool7_kernel<4, 1>") or k.startswith("roi_pool7_kernel<(int)4, (bool)1>"):
                traffic.setdefault("roi_pool+argmax_scan", tot)
    traffic["source"] = (f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (and the rest of {os.path.basename(out)}), "
                         f"{os.path.basename(src)}, config c2, per launch")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    print("wrote", out, len(launches), "launches;", traffic)


if __name__ == "__main__":
    main()
