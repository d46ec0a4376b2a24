#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction counters by source function / line.
usage: python tools/ncu_by_function.py gpurun_out/prof.ncu-rep [object.o]   (needs -lineinfo builds; runs here, no GPU)"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 else "hevc-image-encoder-lite_b200/csrc/hevce_k_g7.o"
core = "hevc-image-encoder-lite_b200/csrc/hevce_core.h"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
ins, cur, inker = [], None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        inker = "hevce_encode_kernel" in l
        continue
    if not inker:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
ithr = hdr.index("Thread Instructions Executed")
data = rows[2:]
print("sass instructions: disasm", len(ins), "report", len(data))
src = open(core).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:HEVCE_HD|template).*?\b(\w+)\s*\([^;]*$", l)
    if m and "HEVCE_HD" in l and not l.strip().startswith("//"):
        funcs.append((i, m.group(1)))


def fn(line):
    name = "?"
    for i, nm in funcs:
        if i <= line:
            name = nm
        else:
            break
    return name


stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
fstall = collections.defaultdict(collections.Counter)
agg, samp, cnt, thr = (collections.Counter() for _ in range(4))
for k in range(min(len(ins), len(data))):
    key = ins[k] or ("?", 0)
    agg[key] += int(data[k][ix]); samp[key] += int(data[k][isamp]); cnt[key] += 1; thr[key] += int(data[k][ithr])
    fk = (key[0], fn(key[1])) if key[0] == "hevce_core.h" else (key[0], "-")
    for ci in stall_cols:
        v = int(data[k][ci] or 0)
        if v:
            fstall[fk][hdr[ci][6:]] += v
tot, ts = sum(agg.values()), sum(samp.values())
print("warp instructions", tot, "samples", ts, "avg active threads %.1f" % (sum(thr.values()) / max(tot, 1)))
fa, fs, fc, ft = (collections.Counter() for _ in range(4))
for (f, ln), v in agg.items():
    k = (f, fn(ln)) if f == "hevce_core.h" else (f, "-")
    fa[k] += v; fs[k] += samp[(f, ln)]; fc[k] += cnt[(f, ln)]; ft[k] += thr[(f, ln)]
print("%inst %samples  #sass  thr/inst  function")
for k, v in fa.most_common(28):
    top = ", ".join(f"{n}:{c / max(fs[k], 1) * 100:.0f}%" for n, c in fstall[k].most_common(4))
    print(f"{v / tot * 100:5.1f} {fs[k] / ts * 100:7.1f} {fc[k]:7d} {ft[k] / max(v, 1):8.1f}  {k[0]}:{k[1]:16s} {top}")
print("top lines")
for k, v in agg.most_common(22):
    s = src[k[1] - 1].strip()[:100] if k[0] == "hevce_core.h" and k[1] <= len(src) else ""
    print(f"{v / tot * 100:5.1f} {samp[k] / ts * 100:6.1f} {cnt[k]:5d} {thr[k] / max(v, 1):5.1f} {k[0]}:{k[1]} {s}")
