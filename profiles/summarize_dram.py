"""Turn the ncu launch list of `tests/gpu_ncu_target.py forward` into the per-image DRAM summary bench.py reports as
`roofline.traffic`.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tests/gpu_ncu_target.py forward
    python profiles/summarize_dram.py gpurun_out/launches.csv <git-hash> profiles/r02_conv_stack_dram.json

The target runs the weight packing and two eager forwards; the LAST image is the tail of the list starting at its
pack_image_c8 launch.  Conv stack = the conv_gemm_kernel launches before the first RoI-pool launch (trunk + RPN 3x3 + RPN
heads); the linear layers' GEMMs come after it."""
import csv
import json
import sys


def rows_of(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = {}
    order = []
    for r in rd:
        i = int(r["ID"])
        if i not in per:
            per[i] = {"name": r["Kernel Name"]}
            order.append(i)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        elif m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}[unit]
        per[i][m] = v
    return [per[i] for i in order]


def main():
    path, git, out = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = rows_of(path)
    starts = [k for k, r in enumerate(rows) if "pack_image_c8" in r["name"]]
    img = rows[starts[-1]:]
    first_roi = next(k for k, r in enumerate(img) if "roi_pool" in r["name"])
    conv = [r for r in img[:first_roi] if "conv_gemm_kernel" in r["name"]]
    dram = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in conv)
    t_all = sum(r["gpu__time_duration.sum"] for r in img)
    t_conv = sum(r["gpu__time_duration.sum"] for r in conv)
    per_kernel = {}
    for r in img:
        k = r["name"].split("(")[0]
        e = per_kernel.setdefault(k, {"launches": 0, "us": 0.0, "dram_mb": 0.0})
        e["launches"] += 1
        e["us"] += r["gpu__time_duration.sum"]
        e["dram_mb"] += (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]) / 1e6
    gemms = [{"kernel": r["name"].split("(")[0].replace("void frcnn::", ""), "us_under_ncu": round(r["gpu__time_duration.sum"], 1),
              "dram_mb": round((r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]) / 1e6, 1)}
             for r in img if "conv_gemm_kernel" in r["name"]]
    res = {
        "conv_stack_dram_bytes_per_step": dram,
        "conv_stack_launches": len(conv),
        "conv_stack_time_share_under_ncu": t_conv / t_all,
        "launches_per_image": len(img),
        "gemm_launches": gemms,        # in launch order = the order of bench.py's `layers` table
        "per_kernel": {k: {"launches": v["launches"], "us": round(v["us"], 1), "dram_mb": round(v["dram_mb"], 1)} for k, v in per_kernel.items()},
        "source": "profiles/r02_launches_dram_forward.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                  "--clock-control none python tests/gpu_ncu_target.py forward; second forward; binary of git %s)" % git,
        "git": git,
    }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
