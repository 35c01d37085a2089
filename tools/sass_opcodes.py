"""SASS opcode histogram of libkd_b200.so per kernel family (`cuobjdump -sass`, runs on the CPU box):
the tracked evidence that the hot kernels are tcgen05 / TMEM / TMA code.

    python tools/sass_opcodes.py > profiles/sass_opcodes.txt

Mnemonics (B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / st
(TMEM), UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tensor load / store (cp.async.bulk.tensor), UBLKCP = cp.async.bulk, UBLKPF = cp.async.bulk.prefetch.L2,
SYNCS = mbarrier ops, LDGMC = multimem.ld_reduce (NVLS in-switch reduction), HMMA / IMMA = legacy mma.sync (must be absent
from the GEMM kernels).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "speech-distill_b200", "libkd_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "UBLKRED", "SYNCS",
       "HMMA", "IMMA", "MUFU.EX2", "MUFU.LG2", "REDUX", "ATOMG", "ATOMS", "RED", "LDG", "STG", "LDS", "STS", "BAR", "MEMBAR",
       "ELECT", "FENCE", "CCTL", "UCGABAR", "ACQBULK", "CGAERRBAR", "LDGMC")


def family(name):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    d = re.sub(r"\(.*", "", d)
    d = d.replace("kd::fused::", "").replace("kd::", "").replace("void ", "").replace("(anonymous namespace)::", "")
    return d[:110]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = kernels.setdefault(family(m.group(1)), collections.Counter())
            cur["__functions__"] += 1
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur is not None:
            op = m.group(1)
            cur["__instructions__"] += 1
            for k in KEY:
                if op == k or op.startswith(k + "."):
                    cur[k] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG") and ".2CTA" in op:
                cur["UTMALDG.2CTA"] += 1
            if op.startswith("UTCBAR") and "MULTICAST" in op:
                cur["UTCBAR.MULTICAST"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: opcode counts per kernel (sm_100a); tools/sass_opcodes.py")
    tot = collections.Counter()
    for name, c in kernels.items():
        keys = [k for k in c if not k.startswith("__")]
        print(f"{name}\n    instructions {c['__instructions__']}" + "".join(f"  {k} {c[k]}" for k in sorted(keys)))
        tot.update(c)
    print("TOTAL  " + "  ".join(f"{k} {v}" for k, v in sorted(tot.items()) if not k.startswith("__")))


if __name__ == "__main__":
    main()
