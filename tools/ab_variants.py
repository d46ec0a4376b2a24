#!/usr/bin/env python
"""Build A/B variants of libhevce_b200 (compile-time switches HEVCE_OPT_*) and time them on the same GPU in one process
each.  usage:  python tools/ab_variants.py build "LPS4=0" "FLUSH=0" ...     (here, no GPU needed)
               python tools/ab_variants.py run [n h w q]                      (on the GPU box)"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hevc-image-encoder-lite_b200", "csrc")
OUT = os.path.join(ROOT, "hevc-image-encoder-lite_b200", "ab")
FLAGS = "-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden".split()

if sys.argv[1] == "build":
    os.makedirs(OUT, exist_ok=True)
    for f in glob.glob(os.path.join(OUT, "*.so")):
        os.remove(f)
    procs = []
    for spec in ["base"] + sys.argv[2:]:
        defs = [] if spec == "base" else ["-DHEVCE_OPT_" + d for d in spec.split(",")]
        name = spec.replace("=", "").replace(",", "_")
        obj = os.path.join(OUT, name + ".o")
        so = os.path.join(OUT, f"libhevce_{name}.so")
        cmd = (f"nvcc {' '.join(FLAGS)} {' '.join(defs)} -c -o {obj} {CSRC}/hevce_cuda.cu && "
               f"nvcc -gencode arch=compute_100a,code=sm_100a -shared -o {so} {obj} {CSRC}/hevce_api.o -Xlinker --exclude-libs,ALL -lpthread && rm {obj}")
        procs.append((spec, subprocess.Popen(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for spec, p in procs:
        out, _ = p.communicate()
        print(spec, "ok" if p.returncode == 0 else "FAILED\n" + out[-2000:])
else:
    args = sys.argv[2:] or ["1036", "64", "64", "2"]
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import workloads as WL
    n, h, w, q = map(int, args)
    np.save("/tmp/ab_inputs.npy", np.stack([WL.config3_image(i)[100:100 + h, 200:200 + w] for i in range(n)]))   # generated once
    code = ("import sys,os,numpy as np; sys.path.insert(0,'hevc-image-encoder-lite_b200'); import hevce_b200 as H;"
            "H.LIB_PATH=sys.argv[1]; q=int(sys.argv[5]); a=np.load('/tmp/ab_inputs.npy'); imgs=[np.ascontiguousarray(x) for x in a];"
            "s=H.Session(0,[i.shape for i in imgs],q); s.upload(imgs); ms=[s.encode() for _ in range(4)]; print(os.path.basename(sys.argv[1]), ' '.join('%.2f'%m for m in ms), flush=True)")
    for rep in range(2):
        for so in sorted(glob.glob(os.path.join(OUT, "*.so"))):
            subprocess.run([sys.executable, "-c", code, so] + args, cwd=ROOT)
