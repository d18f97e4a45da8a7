#!/usr/bin/env python
"""CPU: per-kernel SASS instruction-class counts of the built extension (`cuobjdump -sass`), the evidence that the
hot kernels really use tcgen05 / TMEM / TMA / multimem / cp.async.  Writes profiles/sass_summary.txt and, for the
named hot kernels (FULL below), the complete SASS listing to profiles/sass/<kernel>.sass."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "distribuuuu_b200", "_ext", "b200_kernels.so")
INTERESTING = re.compile(r"^(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTMACMDFLUSH|SYNCS|LDGSTS|LDGDEPBAR|"
                         r"REDG|RED\.|ATOMG|ATOMS|LDGMC|STGMC|REDGMC|MULTIMEM|LD\.E.*\.SYS|ST\.E.*\.SYS|ELECT|BAR\.|DEPBAR)")


# kernels whose complete listing is committed (one instantiation per role)
FULL = [r"conv_gemm_kernel<256, 0, 2>", r"conv_gemm_kernel<256, 0, 1>", r"conv_gemm_kernel<256, 1, 2>", r"conv_gemm_kernel<64, 0, 1>",
        r"conv_gemm_kernel<256, 3, 2>", r"allreduce_sgd_kernel", r"bn_apply_kernel<1, true>", r"bn_bwd_apply_kernel<1, 1>",
        r"bn_bwd_reduce_kernel<1, 1>", r"attn_fwd_kernel", r"attn_bwd_dq_kernel", r"attn_bwd_dkv_kernel", r"dw_fprop", r"ce_topk_kernel",
        r"se_gate_fwd_kernel", r"stem_im2col_kernel<unsigned char>", r"conv3x3_halo_kernel", r"stem_conv_kernel",
        r"bn_relu_pool_fwd_strip_kernel", r"bn_relu_pool_bwd_strip_kernel<true>", r"dw_fprop_fast_kernel<3, 1>"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kernels, name = collections.OrderedDict(), None
    raw = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            raw[name] = []
            continue
        if name:
            raw[name].append(line)
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
    sass_dir = os.path.join(ROOT, "profiles", "sass")
    os.makedirs(sass_dir, exist_ok=True)
    written = 0
    for (mangled, _), pretty in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", pretty).replace("void ", "")
        if any(pat in short for pat in FULL):
            fn = re.sub(r"[^A-Za-z0-9_]+", "_", short.replace("b200::", "")).strip("_") + ".sass"
            body = [l for l in raw[mangled] if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]   # instruction lines, encodings stripped
            body = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in body]
            with open(os.path.join(sass_dir, fn), "w") as f:
                f.write(f"// {pretty}\n// cuobjdump -sass of distribuuuu_b200/_ext/b200_kernels.so (sm_100a), {len(body)} instructions\n")
                f.write("\n".join(body) + "\n")
            written += 1
    print(f"{written} full listings -> {sass_dir}")
    dst = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(dst, "w").write("\n".join(lines) + "\n")
    print(f"{len([k for k in kernels if 'b200' in k])} kernels -> {dst}")


if __name__ == "__main__":
    sys.exit(main())
