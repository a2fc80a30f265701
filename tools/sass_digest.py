#!/usr/bin/env python
"""tools/sass_digest.py -- static SASS digest of the in-tree libswb200.so (no GPU needed): per kernel, the architecture of the cubin, the
register count and the count of the instructions that prove the design claims of DESIGN.md 5 -- TMA bulk-tensor loads (UTMALDG) and their
mbarrier waits (SYNCS), warp shuffles (SHFL), 128-bit global / shared accesses (LDG.E.128, STG.E.128, LDS.128), FP64 conversions (F2F) in the
promoted-arithmetic instantiations, and the absence of tensor-core instructions (HMMA / UTCMMA: the stencils are bandwidth-bound).

    python tools/sass_digest.py > profiles/r2_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "seismicwaves.jl_b200", "libswb200.so")
KEYS = ["UTMALDG", "SYNCS", "SHFL", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "F2F", "DFMA", "FFMA", "HMMA", "UTCMMA", "BAR.SYNC", "PREEXIT", "ACQBULK"]  # PREEXIT / ACQBULK = griddepcontrol.launch_dependents / .wait (programmatic dependent launch)


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"libswb200.so: cubin architectures {arch}")
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[name]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[name][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"{len(kernels)} kernels; columns: instructions, registers, then counts of {', '.join(KEYS)}\n")
    tot = collections.Counter()
    for (mangled, c), dn in zip(kernels.items(), demangled):
        dn = re.sub(r"\(anonymous namespace\)::", "", dn)
        dn = re.sub(r"\((swb::)?\w+<\w+>(, .*)?\)$|\(.*\)$", "", dn).replace("void swb::", "").replace("swb::", "")
        cols = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{dn[:110]:110s} ins={c['_total']:6d} regs={regs.get(mangled, '?'):>3} {cols}")
        tot.update(c)
    print("\nwhole library:", " ".join(f"{k}={tot[k]}" for k in KEYS))
    assert tot["HMMA"] == 0 and tot["UTCMMA"] == 0, "tensor-core instructions in a stencil library?"
    assert tot["UTMALDG"] > 0 and tot["SHFL"] > 0 and tot["LDG.E.128"] > 0


if __name__ == "__main__":
    sys.exit(main())
