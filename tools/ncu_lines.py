"""Dynamic per-source-line attribution: joins an `ncu --page source --csv` export (per-SASS-instruction executed
counts, address order) with nvdisasm's line table of the same kernel in the object file (same order).
    python tools/ncu_lines.py src.csv build/obj/remap_fast.o 'fastILi0ELb0' pixels [min_per_px]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, pat):
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.check_output(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], text=True)
    for sec in re.split(r"\n\s*\.section\s+\.text\.", txt)[1:]:
        if pat not in sec.split(",", 1)[0]:
            continue
        cur, out = None, []
        for line in sec.split("\n"):
            m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', line)
            if m:
                cur = f"{m.group(1)}:{m.group(2)}"
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if m:
                out.append((cur, m.group(1).split(".")[0]))
        return out
    raise SystemExit("kernel not found")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    sl = sass_lines(sys.argv[2], sys.argv[3])
    px = float(sys.argv[4])
    minpp = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
    ci = {h: i for i, h in enumerate(rows[1])}
    body = rows[2:]
    assert len(body) == len(sl), (len(body), len(sl))
    per, ops, smp = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
    for r, (ln, op) in zip(body, sl):
        n = int(r[ci["Instructions Executed"]])
        per[ln] += n
        ops[ln][op] += n
        smp[ln] += int(r[ci["# Samples"]])
    tot, st = sum(per.values()), sum(smp.values())
    print(f"total {tot * 32 / px:.1f} thread-instr/px")
    for ln, n in sorted(per.items(), key=lambda kv: (kv[0].split(':')[0], int(kv[0].split(':')[1]))):
        pp = n * 32 / px
        if pp >= minpp:
            print(f"  {ln:28s} {pp:6.1f}/px  stall {100 * smp[ln] / st:4.1f} %   " +
                  ", ".join(f"{o} {c * 32 / px:.1f}" for o, c in ops[ln].most_common(5)))


if __name__ == "__main__":
    main()
