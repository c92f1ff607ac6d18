"""Per-kernel SASS instruction counts of the built library (cuobjdump -sass; no GPU needed): the evidence that the hot kernels
are tcgen05 / TMEM / TMA code.   python tools/sass_counts.py [lib] > profiles/r02_sass_counts.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "image_restoration_and_enhancement_b200" / "librestoragen.so"
txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
KEYS = [("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("UTCHMMA", r"UTCHMMA(?!\.2CTA)"), ("UTCBAR", r"UTCBAR"), ("LDTM", r"LDTM"), ("STTM", r"STTM"),
        ("UTMALDG", r"UTMALDG"), ("UTMASTG", r"UTMASTG"), ("UTMAREDG", r"UTMAREDG"), ("UBLKCP", r"UBLKCP"), ("SYNCS", r"SYNCS"),
        ("MUFU.EX2", r"MUFU\.EX2"), ("GRIDDEP", r"ACQBULK|GRIDDEPCONTROL|ACQSHMINIT|PREEXIT"), ("USETMAXREG", r"USETMAXREG"),
        ("UCGABAR_ARV", r"UCGABAR_ARV")]
counts, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    for name, pat in KEYS:
        if re.search(pat, line):
            counts[cur][name] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"Per-kernel SASS instruction counts of {lib.relative_to(ROOT)} (cuobjdump -sass, sm_100a build of this tree).")
print("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG / UTMAREDG = TMA tile load / store / reduce,")
print("UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, UCGABAR_ARV = barrier.cluster.arrive (st.shared::cluster compiles to a generic ST.E.128).  Kernels without any of these are the CUDA-core (HBM-bound / glue) kernels.\n")
for (mangled, c), name in zip(counts.items(), demangled):
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    print(name)
    print("    " + ("  ".join(f"{k}={v}" for k, v in c.items()) if c else "(no tensor-core / TMA / mbarrier instructions)"))
