#!/usr/bin/env python
"""DRAM traffic of the separable Gaussian, measured: runs each of the six octave-0 pyramid
filters at n^3 under `ncu` (dram__bytes_read.sum + dram__bytes_write.sum, gpu__time_duration.sum;
--clock-control none) and writes profiles/r02_ncu_blur_traffic.json, which bench.py reports as
`roofline.traffic` (mean over the six launches) and `roofline.traffic_per_filter`.

    python tools/ncu_blur_traffic.py [n=512] [out=profiles/r02_ncu_blur_traffic.json]

The raw ncu CSV of every filter is kept next to the JSON (profiles/r02_ncu_blur_w<width>.csv).
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from bench import gauss_taps, pyramid_filters  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
        "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    out = Path(sys.argv[2]) if len(sys.argv) > 2 else REPO / "profiles" / "r02_ncu_blur_traffic.json"
    out.parent.mkdir(parents=True, exist_ok=True)
    per = []
    for which, sg in enumerate(pyramid_filters()):
        width = len(gauss_taps(sg))
        raw = out.parent / f"r02_ncu_blur_w{width}.csv"
        cmd = ["ncu", "--csv", "--clock-control", "none", "--metrics",
               "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
               "-k", "regex:k_blur_(fused|tma)", "-s", "2", "-c", "1", "--log-file", str(raw),
               sys.executable, str(REPO / "tools" / "run_blur.py"), str(n), str(which), "3"]
        subprocess.run(cmd, check=True, capture_output=True, text=True)
        vals = {}
        rows = [r for r in csv.reader(open(raw)) if len(r) > 5]
        hdr = rows[0]
        im, iu, iv, ik = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), \
            hdr.index("Kernel Name")
        kernel = ""
        for r in rows[1:]:
            vals[r[im]] = float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
            kernel = r[ik]
        rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
        per.append({"width": width, "kernel": kernel, "dram_read_bytes": rd, "dram_write_bytes": wr,
                    "dram_bytes": rd + wr, "algorithmic_bytes": 8.0 * n ** 3,
                    "ratio": round((rd + wr) / (8.0 * n ** 3), 4),
                    "ncu_duration_ms": round(vals["gpu__time_duration.sum"], 4)})
        print(per[-1], flush=True)
    out.write_text(json.dumps({"n": n, "how": "ncu --clock-control none --metrics dram__bytes_read.sum,"
                               "dram__bytes_write.sum,gpu__time_duration.sum, third launch of each filter",
                               "per_filter": per}, indent=1))
    print("wrote", out)


if __name__ == "__main__":
    main()
