"""Builds tests/hostsim/_hostsim.so (g++, CPU).  TEST HARNESS ONLY; see hostsim.cpp."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostsim.so")
SRC = os.path.join(HERE, "hostsim.cpp")
DEPS = [SRC] + [os.path.join(HERE, "../../pymc3_b200/csrc", f)
                for f in ("b2_core.cuh", "b2_models.cuh", "b2_philox.cuh")] + \
       [os.path.join(HERE, "../../include/b200nuts.h")]


def build(force=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return SO
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC, "-o", SO, "-lm"]
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True))
