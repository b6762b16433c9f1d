#!/usr/bin/env python
"""Instruction-mix excerpt of the shipped library (cuobjdump -sass), committed
under profiles/ as the evidence behind the prose claims in DESIGN.md:
mixed-precision adds (FHADD), vector reductions (REDG), 16-byte gathers
(LDG.E.128), bulk copies (UBLKCP) -- and the absence of tensor-core / TMA-tensor
instructions (UTCHMMA/UTCQMMA, UTMALDG, LDTM), because no stage is a dense
contraction.

    python scripts/sass_summary.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuembed_b200", "lib", "libcuembed_b200.so")
WATCH = ["FHADD", "FHADD.BF16", "HADD2", "HFMA2", "FADD", "FFMA", "LDG.E.128", "LDG.E.64", "LDG.E",
         "STG.E.128", "LDS.128", "LDS", "STS", "REDG", "ATOMG", "ATOMS", "RED", "UBLKCP", "LDGSTS",
         "SYNCS", "SHFL", "VOTE", "MATCH", "BAR", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA",
         "UTCBAR", "LDTM", "HMMA", "IMMA", "CCTL", "ERRBAR", "MEMBAR", "NANOSLEEP", "PREFETCH"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True,
                          check=True).stdout
    per_kernel = collections.OrderedDict()
    cur = None
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(2)
            per_kernel[cur][op] += 1
            total[op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per_kernel), capture_output=True,
                              text=True).stdout.splitlines()
    names = dict(zip(per_kernel, demangle))

    def fam(counter, prefix):
        return sum(v for k, v in counter.items() if k == prefix or k.startswith(prefix + "."))

    lines = []
    lines.append(f"# SASS instruction mix of {os.path.relpath(LIB, ROOT)} (sm_100a), "
                 f"{len(per_kernel)} kernels, {sum(total.values())} instructions")
    lines.append("# family counts over the whole library (prefix match on the mnemonic)")
    for w in WATCH:
        lines.append(f"{w:14s} {fam(total, w):8d}")
    lines.append("")
    lines.append("# tensor-core / TMA-tensor / TMEM instructions: "
                 + ", ".join(f"{w}={fam(total, w)}" for w in
                             ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA", "IMMA")))
    lines.append("")
    lines.append("# headline instantiations (C2: fp16 rows of 512 B, int32 indices)")
    want = [r"FwdPoolKernel<__half, 16, int, false, false, 8>", r"FwdHotKernel<__half",
            r"BwdSegReduceKernel<__half, 16, int, false, 8, 0>", r"BwdHot\w*Kernel<__half",
            r"BwdFixupKernel<__half>", r"RadixPassKernel<int, 0, 16>", r"RadixHistKernel<int>",
            r"CompressScanKernel<int>", r"ShardPoolPushKernel<__half, 16, int, false>",
            r"GatherRowsKernel<32, false>"]
    for pat in want:
        for mangled, nice in names.items():
            if re.search(pat, nice):
                c = per_kernel[mangled]
                top = ", ".join(f"{k} {v}" for k, v in c.most_common(12))
                lines.append(f"{nice.split('(')[0]}: {sum(c.values())} instr; {top}")
    text = "\n".join(lines) + "\n"
    if out_path:
        with open(out_path, "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
