"""SASS opcode histogram of the built objects (cuobjdump -sass): which tensor-core / TMA / TMEM instructions each
kernel file contains.  Output: profiles/r2_sass_histogram.txt (committed evidence that the hot kernels are
tcgen05 / TMA code: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads /
stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "foundationpose_cpp_b200", "build")
KEY = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "MUFU",
       "REDG", "ATOMS", "ATOMG", "LDGSTS", "STS", "LDS", "LDG", "STG", "LDC", "BAR", "SHFL", "FFMA", "HFMA2", "IMAD")


def main():
    out = []
    for name in sorted(os.listdir(OBJ)):
        if not name.endswith(".o"):
            continue
        r = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, name)], capture_output=True, text=True)
        if r.returncode != 0:
            continue
        func, per = None, collections.OrderedDict()
        for line in r.stdout.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                func = m.group(1)
                per[func] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
            if m and func:
                per[func][m.group(1)] += 1
        out.append(f"== {name}")
        for f, c in per.items():
            total = sum(c.values())
            if total == 0:
                continue
            short = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
            short = re.sub(r"\(.*", "", short.replace("(anonymous namespace)::", ""))[:110]
            groups = collections.Counter()
            for op, n in c.items():
                base = op.split(".")[0]
                if base in KEY:
                    groups[base] += n
            variants = sorted({op for op in c if op.split(".")[0] in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP")})
            out.append(f"  {short}: {total} instructions; " + ", ".join(f"{k} {groups[k]}" for k in KEY if groups[k]))
            if variants:
                out.append("      variants: " + " ".join(variants))
    text = "\n".join(out) + "\n"
    dst = os.path.join(ROOT, "profiles", "r2_sass_histogram.txt")
    open(dst, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
