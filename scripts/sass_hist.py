"""Static SASS instruction mix of one kernel of the built library: python scripts/sass_hist.py <substring of the demangled name>."""
import collections
import re
import subprocess
import sys

so = "fft_conv_pytorch_b200/libfftconv_b200.so"
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, hist = None, None
res = {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = name if pat in name else None
        if cur:
            res[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(2).split(".")[0]
            full = m.group(2)
            key = full if op in ("LDG", "STG", "LDS", "STS", "LDGSTS") else op
            res[cur][key] += 1
for k, c in res.items():
    tot = sum(c.values())
    print(k, "total", tot)
    print("  ", ", ".join(f"{op} {n}" for op, n in c.most_common(28)))
