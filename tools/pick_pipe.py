"""Reads the output of tools/copy_pipe_ab.py and prints the environment assignments of the fastest
variant (median total of the sparse path), or nothing if copy_pipe=0 is within 1 % of it."""
import re
import sys

best, base = None, None
for line in open(sys.argv[1]):
    m = re.match(r"(copy_pipe=\S+)\s+detect.*total\s+([\d.]+) /\s+([\d.]+)\s+same result", line)
    if not m:
        continue
    spec, med = m.group(1), float(m.group(3))
    if spec == "copy_pipe=0":
        base = med
    if best is None or med < best[1]:
        best = (spec, med)
if best and base and best[0] != "copy_pipe=0" and best[1] < 0.99 * base:
    names = {"copy_pipe": "S3D_COPY_PIPE", "pipe_chunk_kb": "S3D_PIPE_CHUNK_KB", "pipe_slots": "S3D_PIPE_SLOTS"}
    print(" ".join(f"{names[k]}={v}" for k, v in (kv.split("=") for kv in best[0].split("+"))))
