"""Static SASS attribution: instructions per source line (and opcode mix) of one kernel in an object file.

    python tools/sass_lines.py build/obj/remap_fast.o 'fastILi0ELb0' [min_count]

Needs the object to be compiled with -lineinfo (the Makefile does).  Uses cuobjdump -xelf + nvdisasm --print-line-info.
Static counts: every instruction once, whatever its trip count — a guide for where the instructions ARE, the dynamic
numbers come from ncu (profiles/)."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
    min_count = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", obj], cwd=d, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.check_output(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], text=True)
    sections = re.split(r"\n\s*\.section\s+\.text\.", txt)
    for sec in sections[1:]:
        name = sec.split(",", 1)[0]
        if pat not in name:
            continue
        cur, cnt, ops, allops = None, collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
        for line in sec.split("\n"):
            m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', line)
            if m:
                cur = (m.group(1), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if m and cur:
                op = m.group(1).split(".")[0]
                cnt[cur] += 1
                ops[cur][op] += 1
                allops[op] += 1
        print(f"== {name[:100]}\n   total static instructions: {sum(cnt.values())}")
        print("   opcode mix:", ", ".join(f"{k} {v}" for k, v in allops.most_common(24)))
        for k, v in sorted(cnt.items()):
            if v >= min_count:
                print(f"   {k[0]}:{k[1]:<5d} {v:5d}  ", ", ".join(f"{o} {n}" for o, n in ops[k].most_common(5)))


if __name__ == "__main__":
    main()
