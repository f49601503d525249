#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove which hardware paths a kernel uses (tcgen05 = UTC*MMA / LDTM / UTCBAR,
TMA / bulk copies = UTMALDG / UBLKCP, per-thread async copies = LDGSTS, packed fp32 = FFMA2 / FMUL2 / FADD2) and of the
waterfall loops (BRA.U.ANY) that a non-elected issue guard produces.  python tools/sass_counts.py > profiles/r02_sass_tc.txt"""
import collections, glob, os, re, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "ELECT", "BRA.U.ANY", "R2UR",
       "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "HMMA", "ACQBULK", "UGETNEXTWORKID"]
objs = sorted(glob.glob(os.path.join(REPO, "birdnet-go_b200", "build", "*.o")))
print("# cuobjdump -sass of birdnet-go_b200/build/*.o (sm_100a): opcode counts per kernel; total = all SASS instructions of the kernel")
for o in objs:
    out = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    cur, per = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur); cur = re.sub(r"\(.*", "", cur).replace("bnb::", "").replace("void ", "")
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["total"] += 1
            for p in PAT:
                if op == p or op.startswith(p + ".") or (p == "BRA.U.ANY" and op.startswith("BRA.U.ANY")):
                    per[cur][p] += 1
    for k, c in per.items():
        if c["total"] == 0:
            continue
        print("%-34s %-44s total %5d  %s" % (os.path.basename(o), k[:44], c["total"], " ".join("%s=%d" % (p, c[p]) for p in PAT if c[p])))
