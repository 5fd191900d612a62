#!/usr/bin/env python
"""Instruction-class digest of libpic_latent.so per kernel (cuobjdump -sass): what proves the sm_100a paths --
packed f32 (FFMA2 / FMUL2 / FADD2), cp.async (LDGSTS), 1-D TMA bulk copies (UBLKCP) with mbarriers (SYNCS), cluster /
distributed-shared-memory ops, shared atomics, MUFU.  usage: python scripts/sass_digest.py [lib.so] > profiles/..."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "efficient-pic-with-variance-aware-masking_b200", "libpic_latent.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
CLASSES = [("packed_f32", r"^(FFMA2|FMUL2|FADD2)"), ("fp32", r"^(FFMA|FMUL|FADD|FSETP|FSET|FMNMX|FSEL)"), ("mufu", r"^MUFU"),
           ("ldg", r"^LDG"), ("stg", r"^STG"), ("lds", r"^LDS"), ("sts", r"^STS"), ("cp_async(LDGSTS)", r"^LDGSTS"),
           ("tma_bulk(UBLKCP)", r"^UBLKCP"), ("tma_tensor(UTMA*)", r"^UTMA"), ("mbarrier(SYNCS)", r"^SYNCS"),
           ("shared_atomic(ATOMS)", r"^ATOMS"), ("global_atomic(ATOMG/RED)", r"^(ATOMG|RED|ATOM\b)"), ("barrier(BAR)", r"^BAR"),
           ("cluster(UCGABAR/ cluster ld)", r"^(UCGABAR|CGAERRBAR|LDS\.CLUSTER|MAPA)"), ("shuffle/vote/redux", r"^(SHFL|VOTE|REDUX|MATCH)"),
           ("tensor(HMMA/UTC*MMA)", r"^(HMMA|UTC|IMMA|DMMA)")]
kern, counts, arch = None, collections.OrderedDict(), set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("pic::", "")
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for name, pat in CLASSES:
            if re.match(pat, op):
                counts[kern][name] += 1
print(f"# SASS digest of {os.path.relpath(lib, ROOT)}  (architectures in the fatbin: {', '.join(sorted(arch))})")
cols = ["total"] + [c for c, _ in CLASSES]
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    shown = "  ".join(f"{n}={c[n]}" for n in cols if c[n])
    print(f"{k[:110]}\n    {shown}")
print("\nALL KERNELS\n    " + "  ".join(f"{n}={tot[n]}" for n in cols if tot[n]))
