#!/bin/bash
# The Blackwell proof: per kernel of libpp_b200.so, counts of tcgen05 MMA (UTCHMMA/UTCQMMA...), TMEM loads (LDTM),
# TMA tensor / bulk copies (UTMALDG, UTMASTG, UBLKCP), mbarrier ops (SYNCS), registers and static shared memory.
# Runs without a GPU: bash scripts/sass_summary.sh > profiles/sass_summary.txt
cd "$(dirname "$0")/.."
so=3d-object-detection_b200/libpp_b200.so
python - "$so" <<'PY'
import collections, re, subprocess, sys
so = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
ops = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "ELECT", "HMMA", "FFMA2", "FADD2"]
counts = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[name]["total"] += 1
        if op in ops:
            counts[name][op] += 1
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("%-64s %6s %6s %6s " % ("kernel (sm_100a SASS of libpp_b200.so)", "instr", "regs", "smem") + " ".join("%7s" % o for o in ops))
for (mangled, c), nice in zip(counts.items(), demangle):
    nice = re.sub(r"\(.*", "", nice).replace("pp::", "")
    r, s = usage.get(mangled, (0, 0))
    print("%-64s %6d %6d %6d " % (nice[:64], c["total"], r, s) + " ".join("%7d" % c[o] for o in ops))
PY
