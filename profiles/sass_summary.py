"""SASS evidence for the kernels of libb200nuts.so: mnemonic counts per kernel and three excerpts of the shipped
tcgen05 likelihood kernel.  Usage: python profiles/sass_summary.py > profiles/sass_tcgen05_r2.txt  (needs cuobjdump)."""
import collections
import os
import re
import subprocess
import sys

SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pymc3_b200", "libb200nuts.so")
txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kernels, name = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = []
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        kernels[name].append(line.split("/* 0x")[0].rstrip())

WATCH = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "UTCATOMSWS", "MUFU", "FFMA2", "FADD2", "FMUL2", "FFMA",
         "DADD", "DFMA", "LDGSTS", "BAR"]
print("SASS evidence for the kernels of libb200nuts.so (cuobjdump -sass of the in-tree build, sm_100a; round 2, shipped build;\n"
      "regenerate with profiles/sass_summary.py).  Mnemonic counts per kernel: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,\n"
      "UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D bulk copies: X is stored pre-swizzled, so there is no tensor map and no\n"
      "UTMALDG), SYNCS = mbarrier ops, UTCATOMSWS = TMEM allocation, FFMA2 / FADD2 / FMUL2 = packed fp32 math, LDGSTS = cp.async.\n")
want = ("k_glm_tc_main", "k_glm_tcw_main", "k_glm_tc_post", "k_hier_slab", "k_persistent_block", "k_dense")
for k, ins in kernels.items():
    if not any(w in k for w in want):
        continue
    cnt = collections.Counter()
    for l in ins:
        op = re.sub(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?", "", l).split()[0].rstrip(";")
        base = op.split(".")[0]
        if base in WATCH:
            cnt[base] += 1
    print(k)
    print("    %d instructions; %s" % (len(ins), ", ".join("%s %d" % (w, cnt[w]) for w in WATCH if cnt[w])))


def excerpt(kernel, pattern, before, after, title):
    ins = kernels[kernel]
    idx = next(i for i, l in enumerate(ins) if re.search(pattern, l))
    print("\n" + title)
    print("\n".join(ins[max(0, idx - before):idx + after]))


main = next(k for k in kernels if "k_glm_tc_mainILi4ELi0ELb0" in k)
excerpt(main, r"UTCHMMA", 6, 14, "Excerpt, %s: the first GEMM1 MMAs of the issue loop (A operand from TMEM, B = shared-memory "
        "descriptor in uniform registers)" % main)
excerpt(main, r"LDTM", 2, 16, "Excerpt: epilogue, TMEM load of S and the first ops after it")
excerpt(main, r"UBLKCP", 6, 18, "Excerpt: producer, bulk copies of one pipeline stage (X tile 32 KB with an L2 evict-first policy, y block, "
        "eta_ref block) onto one mbarrier")
