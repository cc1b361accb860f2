"""Build libwisp_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m wisp_mdanalysis_implementation_b200.build [--force]

The shared library sits next to this file so it travels with the repository snapshot to the
GPU box; it is git-ignored.  Sources are recompiled only when newer than their objects.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwisp_b200.so")
SOURCES = ["capi.cu", "k1_com.cu", "k1_stream.cu", "k2_gram.cu", "k2_gram_i8.cu", "k3_contact.cu", "k4_epilogue.cu",
           "k5_apsp.cu", "k6_paths.cu", "k7_align.cu", "k8_lmi.cu", "textio.cu", "pipeline.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
# tuning experiments: WISP_NVCC_DEFS="-DWISP_WS_FC=16 -DWISP_WS_STAGES=2" python -m ...build --force
NVCC_FLAGS += os.environ.get("WISP_NVCC_DEFS", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".cuh")]
    headers.append(os.path.join(HERE, "..", "include", "wisp_b200.h"))
    hdr_m = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _newer(s, o) or hdr_m > os.path.getmtime(o):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
            return o, True
        return o, False

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if force or any(ch for _, ch in res) or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
