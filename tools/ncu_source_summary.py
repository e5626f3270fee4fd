"""Summarises an `ncu --page source --csv` export (SASS view): executed warp-instructions per opcode and per pipe
class, stall samples per opcode, shared-memory wavefronts.  Usage:
    ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_summary.py src.csv [pixels]"""
import collections
import csv
import re
import sys

FMA = {"FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "IMAD", "HFMA2", "I2FP"}  # fma pipe (B300_MICROARCH: FFMA/FMUL/IMAD/HFMA2)
XU = {"MUFU", "F2I", "I2F", "FRND", "F2F", "POPC", "FLO", "BREV"}
LSU = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "LDSM", "RED", "ATOMG"}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    px = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    ex, samp, wav, ideal = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[2:]:
        m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", r[ci["Source"]])
        if not m:
            continue
        op = m.group(1)
        ex[op] += int(r[ci["Instructions Executed"]])
        samp[op] += int(r[ci["# Samples"]])
        wav[op] += int(r[ci["L1 Wavefronts Shared"]] or 0)
        ideal[op] += int(r[ci["L1 Wavefronts Shared Ideal"]] or 0)
    tot, stot = sum(ex.values()), sum(samp.values())
    cls = collections.Counter()
    for op, n in ex.items():
        cls["fma" if op in FMA else "xu" if op in XU else "lsu" if op in LSU else "alu/other"] += n
    per = (lambda n: f" = {n * 32 / px:7.1f} thread-instr/px") if px else (lambda n: "")
    print(f"executed warp-instructions: {tot}{per(tot)}")
    for k, n in cls.most_common():
        print(f"  pipe class {k:10s} {n:10d} {100 * n / tot:5.1f} %{per(n)}")
    print("  opcode      executed   share  stall-samples  smem wavefronts (ideal)")
    for op, n in ex.most_common(28):
        print(f"  {op:10s} {n:10d} {100 * n / tot:5.1f} %   {100 * samp[op] / max(stot, 1):5.1f} %      {wav[op]:9d} ({ideal[op]})")


if __name__ == "__main__":
    main()
