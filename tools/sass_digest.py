"""profiles/r02_sass_digest.txt: per kernel of libveonlift.so, the number of SASS instructions and
of the mnemonics that prove (or rule out) the Blackwell paths.  python tools/sass_digest.py > FILE"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "veon_b200/libveonlift.so"
COLS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "SYNCS",
        "USETMAXREG", "REDUX", "ATOM", "RED"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = {}
for line in subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout.splitlines():
    pass
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_n"] += 1
        counts[cur][op.split(".")[0]] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass veon_b200/libveonlift.so: per kernel, number of SASS instructions and of the")
print("# mnemonics that prove (or rule out) the Blackwell paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma,")
print("# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA (tensor / bulk copies), LDGSTS = cp.async,")
print("# SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, HMMA = legacy mma.sync (none expected)")
print("kernel,sass_instructions," + ",".join(COLS))
for (mangled, c), name in sorted(zip(counts.items(), demangled), key=lambda t: t[1]):
    short = re.sub(r"\(.*", "", name)
    print(f'"{short}",{c["_n"]},' + ",".join(str(c[k]) for k in COLS))
