"""Per-source-line warp-stall samples of one kernel from an ncu report (needs -lineinfo builds).
usage: python tools/ncu_lines.py report.ncu-rep object.o demangled_substring mangled_substring source.cu [top]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, obj, kern, mangled, src_path = sys.argv[1:6]
    top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    s = [x for x in secs if kern in x["name"]][0]
    subprocess.run(f"cd /tmp && cuobjdump -xelf all {obj} > /dev/null 2>&1", shell=True)
    cub = subprocess.run("ls -t /tmp/*.cubin | head -1", shell=True, capture_output=True, text=True).stdout.strip()
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l]
    start = starts[0]
    end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//---")), len(dis))
    cur_line, a2l = None, {}
    for l in dis[start:end]:
        m = re.search(r'//## File "(.*)", line (\d+)', l)
        if m:
            f = m.group(1).split("/")[-1]
            cur_line = int(m.group(2)) if f == src_path.split("/")[-1] else (f, int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            a2l[int(m.group(1), 16)] = cur_line
    h = s["hdr"]
    ix = {n: i for i, n in enumerate(h)}
    base = int(s["rows"][0][ix["Address"]], 16)
    by = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    stall_names = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    ops = collections.Counter()
    for r in s["rows"]:
        ln = a2l.get(int(r[ix["Address"]], 16) - base, "?")
        e = by[ln]
        e[0] += int(r[ix["# Samples"]] or 0)
        ne = int(r[ix["Instructions Executed"]] or 0)
        e[1] += ne
        for n in stall_names:
            e[2][n] += int(r[ix[n]] or 0)
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
        ops[m.group(2) if m else "?"] += ne
    tot = sum(e[0] for e in by.values()) or 1
    toti = sum(e[1] for e in by.values()) or 1
    src = open(src_path).read().split("\n")
    print(f"kernel {s['name'][:80]}  samples {tot}  warp-instructions {toti}")
    print("opcodes:", ", ".join(f"{o} {100 * c / toti:.1f}%" for o, c in ops.most_common(14)))
    agg = collections.Counter()
    for e in by.values():
        agg.update(e[2])
    print("stalls:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in agg.most_common(8)))
    for ln, e in sorted(by.items(), key=lambda x: -x[1][0])[:top]:
        text = src[ln - 1].strip()[:80] if isinstance(ln, int) else str(ln)
        st = ", ".join(f"{k[6:]} {v}" for k, v in e[2].most_common(2))
        print(f"{str(ln):>26s} smp {100 * e[0] / tot:5.1f}% inst {100 * e[1] / toti:5.1f}% [{st}] | {text}")


if __name__ == "__main__":
    main()
