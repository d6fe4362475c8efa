"""SASS instruction histogram of the tensor-core kernels in libngm_b200.so (no GPU needed):
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
Counts the mnemonics that prove the Blackwell-native path (B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UBLKCP (cp.async.bulk, TMA engine), UTCBAR (tcgen05.commit), SYNCS (mbarrier), and the absence of the
legacy HMMA / HGMMA paths."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "neural_graph_mapping_b200", "libngm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOM", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "HGMMA", "FFMA",
       "HFMA2", "F2FP", "MUFU", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "SHFL", "BAR", "NANOSLEEP", "ELECT", "LDL", "STL")
want = [w for w in sys.argv[1:]] or ["tc_kernel", "bwd_kernel", "composite_staged_kernel", "sample_rays_kernel"]
fn = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ngm::", "").replace("ngm::", "")
        fn = name if any(w in name for w in want) else None
        if fn:
            hist[fn] = collections.Counter()
        continue
    if fn:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            hist[fn]["TOTAL"] += 1
            for k in KEY:
                if op.startswith(k):
                    hist[fn][k] += 1
print(f"{'kernel':44s} " + " ".join(f"{k:>8s}" for k in ("TOTAL",) + KEY))
for fn, h in hist.items():
    print(f"{fn[:44]:44s} " + " ".join(f"{h.get(k, 0):8d}" for k in ("TOTAL",) + KEY))
