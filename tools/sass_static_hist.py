#!/usr/bin/env python
"""Static SASS histogram of the shipped library, per kernel family (no GPU needed):

    python tools/sass_static_hist.py [path/to/libradiocore_b200.so] > profiles/r02_sass_histogram.txt

Uses `cuobjdump -sass`.  Reports, for the whole library and for each kernel family, the static
instruction counts of the mnemonics that identify the Blackwell-specific paths: UTMALDG (TMA tile
loads), SYNCS (mbarrier), packed fp32 (FFMA2 / FADD2 / FMUL2), DFMA (fp64 FIRs / twiddle
recurrences), and that no tensor-core opcode is present (the path is bandwidth-bound butterflies)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "radio-core_b200", "radiocore", "_native", "libradiocore_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
FAMILIES = [("v3_first_kernel", "first FFT pass (fused loaders: TMA tile / tuner gather / discriminator)"),
            ("v3_later_kernel", "later FFT passes (fused stores: angle, window, lmr, peer scatter)"),
            ("fft_pass_kernel", "generic shared-memory pass (sizes without a register-radix split)"),
            ("ew_kernel", "elementwise functors (spectral resample / taper / Hilbert / stereo, sub-band combine)"),
            ("epi_fir_kernel", "de-emphasis FIR + block mean + clip"),
            ("filtfilt_fold_kernel", "folded pilot filter (fp32 pairs)"), ("filtfilt_kernel", "zero-phase FIR, exact fp64")]
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMUL",
         "LDS", "STS", "LDG", "STG", "BAR", "MUFU", "F2F", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "ATOMG", "RED"]
arch = set(re.findall(r"arch = (sm_\w+)", out))
kern, counts, nkern = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern = next((f for f, _ in FAMILIES if f in name), "other")
        nkern[kern] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        op = m.group(1)
        counts[kern][op] += 1
        counts["ALL"][op] += 1
print(f"library: {os.path.relpath(lib, ROOT)}   architectures in the fatbin: {sorted(arch)}")
print(f"kernels (template instantiations): {sum(nkern.values())}")
for fam, what in [("ALL", "whole library")] + FAMILIES + [("other", "anything else")]:
    c = counts.get(fam)
    if not c:
        continue
    tot = sum(c.values())
    print(f"\n== {fam} -- {what}; {nkern.get(fam, sum(nkern.values()))} kernels, {tot} static instructions")
    print("   " + "  ".join(f"{op}={c[op]}" for op in WATCH if c[op]))
    top = ", ".join(f"{op} {100.0 * n / tot:.1f}%" for op, n in c.most_common(8))
    print("   top: " + top)
tc = sum(counts["ALL"][o] for o in ("HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM"))
print(f"\ntensor-core / TMEM opcodes in the library: {tc} (none expected: no dense contraction on this path)")
