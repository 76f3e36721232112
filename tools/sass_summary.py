#!/usr/bin/env python
"""Summarise the SASS of the built library for profiles/: per kernel, counts of the mnemonics that
prove the Blackwell-native path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA
tensor load/store, UBLKCP = bulk copy, SYNCS = mbarrier, UTCBAR = tcgen05.commit) plus the MMA issue
sequence of one kernel.   python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = [os.path.join(ROOT, "tts_king_b200", "build", f) for f in ("conv_tc.o", "conv_tc2.o", "conv_pair_tc.o", "conv_pair_fold.o",
                                                                    "conv_narrow.o", "tail.o", "conv_ffma.o")]
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UTCATOM", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "R2UR", "ELECT"]
show = "conv_tc_kernelILi128ELi64ELi2ELb0"
for o in objs:
    if not os.path.exists(o):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    cur, counts, lines = None, collections.OrderedDict(), {}
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            lines[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if cur and m:
            counts[cur][m.group(1).split(".")[0]] += 1
            lines[cur].append(ln.split("/*")[1].split("*/")[1].rstrip(" ;") if "*/" in ln else ln)
    print(f"== {os.path.basename(o)}")
    for fn, c in counts.items():
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        tot = sum(c.values())
        print(f"{dem[:110]}\n    instrs={tot} " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    for fn in lines:
        if show in fn:
            idx = [i for i, l in enumerate(lines[fn]) if "UTCHMMA" in l]
            if idx:
                print(f"\n-- MMA issue sequence of {show} (SASS lines {idx[0]-24}..{idx[-1]+6}):")
                for l in lines[fn][max(0, idx[0] - 24): idx[-1] + 6]:
                    print("   ", l.strip()[:120])
