"""SASS mnemonic counts per kernel of the in-tree objects (evidence that the hot kernels are Blackwell-native):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit,
UTCATOMSWS = tcgen05.alloc, HMMA = mma.sync (token kernels).  Usage: python profiles/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections, os, re, subprocess
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hdenseformer_b200", "lib")
PAT = re.compile(r"\b(UTCHMMA|LDTM|STTM|UTMALDG|UTCBAR|UTCATOMSWS|HMMA\.[0-9A-Z.]+|SYNCS\.[A-Z0-9.]+)")
for obj in ("tc_conv.o", "tc_conv_ws.o", "tc_wgrad_ws.o", "tc_convt.o", "stem_tc.o", "patch_tc.o", "tok_tc.o"):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
    fn, cnt = None, collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::", "", d).split("(")[0]
            continue
        m = PAT.search(line)
        if m and fn:
            cnt[(fn, m.group(1))] += 1
    for (f, k), v in sorted(cnt.items()):
        print(f"{obj:16s} {f:40s} {k:28s} {v}")
