"""Builds pymc3_b200/libb200nuts.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libb200nuts.so")
SOURCES = ["b2_engine.cu", "b2_glm_simt.cu", "b2_hier.cu", "b2_glm_tc.cu", "b2_glm_tcw.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _deps():
    out = [os.path.join(HERE, "..", "include", "b200nuts.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh")):
            out.append(os.path.join(CSRC, f))
    return out


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in _deps()):
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", os.path.join(CSRC, src), "-o", obj]
        cmd += os.environ.get("B2_NVCC_EXTRA", "").split()          # e.g. -DB2_TC_NOTRAP (debugging the tcgen05 pipelines)
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs + ["-lcuda"], check=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
