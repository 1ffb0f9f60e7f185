#!/usr/bin/env python
"""Evidence that the hot kernels are Blackwell-native: per kernel of libsimhand_b200.so the counts of the SASS mnemonics behind
tcgen05.mma (UTC*MMA), tcgen05.ld/st (LDTM / STTM), cp.async.bulk (UBLKCP), TMA tensor loads (UTMALDG), mbarrier transactions
(SYNCS) and the MUFU flavours, plus a short listing around the first MMA of the sweeps and of the head's GEMM 1.
    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "simhand_b200", "lib", "libsimhand_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "MUFU.SQRT", "MUFU.RSQ", "MUFU.EX2",
        "MUFU.RCP", "FFMA2", "FADD2", "FMUL2", "RED", "ATOM", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]  # noqa: E731
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            kernels[cur].append(ln)
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)\n")
    print(f"{'kernel':70s} " + " ".join(f"{k:>9s}" for k in KEYS))
    for name, lines in kernels.items():
        text = "\n".join(lines)
        counts = [len(re.findall(r"\b" + re.escape(k), text)) for k in KEYS]
        if sum(counts[:8]) == 0 and "mpjpe" not in name and "altdist" not in name:
            continue
        print(f"{demangle(name)[:70]:70s} " + " ".join(f"{c:9d}" for c in counts))
    for want, title in (("sweep_tc_kernelILb1ELb1ELb1ELb0", "backward sweep (bf16 operands, 16-bit tiles): around the first value MMA"),
                        ("head_gemm1_kernel", "projection head GEMM 1: the MMA issue loop")):
        for name, lines in kernels.items():
            if want in name:
                idx = [i for i, ln in enumerate(lines) if "UTCHMMA" in ln]
                if idx:
                    print(f"\n# {title}\n# {demangle(name)}")
                    lo, hi = max(0, idx[0] - 12), min(len(lines), idx[0] + 14)
                    for ln in lines[lo:hi]:
                        print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln.rstrip()))
                break


if __name__ == "__main__":
    sys.exit(main())
