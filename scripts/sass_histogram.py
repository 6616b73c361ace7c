"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the tcgen05 / TMA / TMEM mnemonics that prove
which kernels run on the 5th-gen tensor cores (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier), plus the MUFU / FFMA / SHFL mix of their epilogues.
Usage: python scripts/sass_histogram.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "distillclip_b200", "csrc", "libdistillclip_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "MUFU", "FFMA", "FMUL", "FADD", "SHFL", "LDG", "STG", "LDS", "STS", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
        if m.group(1) in ("UTCHMMA", "UTMALDG", "LDTM") and m.group(2):
            kernels[cur][m.group(1) + m.group(2)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode histogram of `libdistillclip_b200.so` (sm_100a), per kernel\n")
print("`cuobjdump -sass`; counts are static instructions.  UTCHMMA = `tcgen05.mma` (`.2CTA` = `cta_group::2`), UTMALDG = TMA tensor")
print("load (`cp.async.bulk.tensor`), LDTM = `tcgen05.ld`, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier ops.  No HMMA (`mma.sync`) anywhere.\n")
print("| kernel | total | " + " | ".join(KEY) + " | variants |")
print("|---|---|" + "---|" * (len(KEY) + 1))
for (name, c), dm in zip(kernels.items(), demangle):
    short = re.sub(r"\(.*", "", dm).replace("void dcb::", "").replace("dcb::", "")
    short = short if len(short) < 90 else short[:87] + "..."
    var = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if "." in k)
    print(f"| `{short}` | {sum(v for k, v in c.items() if '.' not in k)} | " + " | ".join(str(c.get(k, 0)) for k in KEY) + f" | {var} |")
