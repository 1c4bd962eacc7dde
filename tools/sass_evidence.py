"""Counts of the SASS mnemonics that identify Blackwell-native code (B200_PROFILING.md, "What proves a Blackwell-native
kernel") in every kernel of the built library: UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA
engine), UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, SYNCS = mbarrier, F*2 = packed fp32.  No GPU needed.

    python tools/sass_evidence.py > profiles/r1_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cough_detector_b200", "lib", "libcough_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|UTMASTG|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|UCGABAR_\w+|SYNCS|HMMA|"
                 r"FADD2|FMUL2|FFMA2|MUFU\.LG2)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cnt[cur] = collections.Counter()
            continue
        if cur:
            for t in PAT.findall(ln):
                cnt[cur][t] += 1
    print(subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.strip())
    for name, c in cnt.items():
        if c:
            print(f"{name}: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
