#!/usr/bin/env python
"""CPU: per-kernel SASS instruction-class counts of the built extension (`cuobjdump -sass`), the evidence that the
hot kernels really use tcgen05 / TMEM / TMA / multimem / cp.async.  Writes profiles/sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "distribuuuu_b200", "_ext", "b200_kernels.so")
INTERESTING = re.compile(r"^(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTMACMDFLUSH|SYNCS|LDGSTS|LDGDEPBAR|"
                         r"REDG|RED\.|ATOMG|ATOMS|LDGMC|STGMC|REDGMC|MULTIMEM|LD\.E.*\.SYS|ST\.E.*\.SYS|ELECT|BAR\.|DEPBAR)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
        if m and name:
            ins = re.sub(r"^@!?U?P\d\s+", "", m.group(1).strip())
            kernels[name].append(ins.split()[0] if ins else "")
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    lines = ["# SASS evidence (tools/sass_summary.py over distribuuuu_b200/_ext/b200_kernels.so, sm_100a), per kernel: instruction-class counts",
             "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc, UTMALDG / UTMASTG = TMA load / store",
             "# (IM2COL = im2col mode), SYNCS = mbarrier ops, LDGSTS / LDGDEPBAR = cp.async, REDG = red.global (F32x4 = vector),",
             "# LDGMC...HPADD = multimem.ld_reduce (multimem.st compiles to a plain STG.E.128 on the multicast address and is not listed)", ""]
    for (mangled, ins), pretty in zip(kernels.items(), demangle):
        if "b200" not in mangled:
            continue
        c = collections.Counter(i for i in ins if INTERESTING.match(i))
        short = re.sub(r"\(.*", "", pretty).replace("void ", "")
        lines.append(f"## {short}  ({len(ins)} instructions)")
        lines.append("   " + (", ".join(f"{k} x{v}" for k, v in sorted(c.items())) or "(no tensor / TMA / async / atomic instructions)"))
    dst = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(dst, "w").write("\n".join(lines) + "\n")
    print(f"{len([k for k in kernels if 'b200' in k])} kernels -> {dst}")


if __name__ == "__main__":
    sys.exit(main())
