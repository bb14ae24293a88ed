#!/usr/bin/env python
"""Turn the round-2 ncu captures in gpurun_out/ into the text summaries kept under profiles/
(the .ncu-rep files themselves are scratch).  Usage: python tools/r2_profiles.py"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
G, P = REPO / "gpurun_out", REPO / "profiles"

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def raw_summary(rep, title, cmd):
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [title, "command: " + cmd, ""]
    for i, h in enumerate(hdr):
        if h == "Kernel Name":
            lines.append(f"kernel: {vals[i]}")
        if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.2):
            lines.append(f"{h:100s} {vals[i]:>18s} {units[i]}")
    return lines


def opcode_hist(rep):
    rows = ncu_csv(rep, "source")
    hdr, data = rows[1], rows[2:]
    iS, iE, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    c, t = collections.Counter(), collections.Counter()
    for r in data:
        w = r[iS].strip().split()
        if not w:
            continue
        op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
        c[op] += int(r[iE])
        t[op] += int(r[iT])
    tot = sum(c.values())
    lines = ["", "executed warp-instructions by opcode (source page): share, average active lanes"]
    for op, n in c.most_common(24):
        lines.append(f"  {op:10s} {100 * n / tot:5.1f} %   {t[op] / max(n, 1):5.1f} lanes")
    lines.append(f"  total {tot / 1e9:.3f} G warp-instructions, {sum(t.values()) / max(tot, 1):.1f} lanes on average")
    return lines


def sass_hist(obj, pattern):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    cur, hist = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1) if re.search(pattern, m.group(1)) else None
            if cur:
                hist[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", ln)
        if cur and m:
            w = m.group(1).split()
            op = w[1] if w[0].startswith("@") else w[0]
            hist[cur][op.split(".")[0] if not op.startswith(("UTMA", "SYNCS", "ATOMS", "RED", "FFMA2", "FMUL2")) else op] += 1
    return hist


def main():
    P.mkdir(exist_ok=True)
    jobs = [("r2_desc3c.ncu-rep", "r02_ncu_desc3.txt", "k_descriptor3<4,true> (cell-owner lanes), first 4096-keypoint chunk of a 512^3 bench step",
             "ncu --set full --clock-control none --import-source on -k regex:k_descriptor3 -c 1 python bench.py --steps 1 --warmup 3 ..."),
            ("r2_orient_group.ncu-rep", "r02_ncu_orient_group.txt", "k_orient_group (one batch in flight; 256^3 volume, 23.4k candidates)",
             "ncu --set full --clock-control none --import-source on -k regex:k_orient_group -c 1 python tools/run_desc.py 256"),
            ("r2_blur_f0.ncu-rep", "r02_ncu_blur_w5.txt", "k_blur_fused<2> (w = 5) at 512^3", "ncu --set full ... python tools/run_blur.py 512 0 3"),
            ("r2_blur_f2.ncu-rep", "r02_ncu_blur_w9.txt", "k_blur_fused<4> (w = 9) at 512^3", "ncu --set full ... python tools/run_blur.py 512 2 3"),
            ("r2_blur_tma_f0.ncu-rep", "r02_ncu_blur_tma_w5.txt", "k_blur_tma<2,4> (w = 5, TMA fill, 64 x 64 tile) at 512^3",
             "ncu --set full --clock-control none --import-source on -k regex:k_blur_tma -s 2 -c 1 python tools/run_blur.py 512 0 3"),
            ("r2_blur_tma_f5.ncu-rep", "r02_ncu_blur_tma_w17.txt", "k_blur_tma<8,2> (w = 17, TMA fill, 64 x 32 tile) at 512^3",
             "ncu --set full --clock-control none --import-source on -k regex:k_blur_tma -s 2 -c 1 python tools/run_blur.py 512 5 3")]
    for rep, out, title, cmd in jobs:
        if not (G / rep).exists():
            print("missing", rep)
            continue
        lines = raw_summary(G / rep, title, cmd) + opcode_hist(G / rep)
        (P / out).write_text("\n".join(lines) + "\n")
        print("wrote", out)
    lines = ["SASS opcode histograms of the product kernels (cuobjdump -sass sift3d_b200/lib/*.o; static counts)", ""]
    for obj, pat in (("keypoint.o", r"k_descriptor3ILi4ELb1|k_orient_group|k_gradient"), ("blur_fused.o", r"k_blur_fusedILi(2|8)ELb1|k_blur_tmaILi(2ELi4|8ELi2)")):
        for fn, h in sass_hist(REPO / "sift3d_b200" / "lib" / obj, pat).items():
            lines.append(f"{obj}: {fn}")
            lines.append("  " + ", ".join(f"{op} {n}" for op, n in h.most_common(28)))
            lines.append(f"  TMA / mbarrier opcodes (UTMALDG, UBLKCP, SYNCS): "
                         f"{sum(n for op, n in h.items() if op.startswith(('UTMA', 'UBLKCP', 'SYNCS')))}")
            lines.append("")
    (P / "r02_sass_opcodes.txt").write_text("\n".join(lines))
    print("wrote r02_sass_opcodes.txt")


if __name__ == "__main__":
    main()
