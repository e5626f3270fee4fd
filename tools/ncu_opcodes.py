"""Dynamic opcode histogram of a kernel from an ncu report (--set full --import-source on), per unit of work.
usage: python tools/ncu_opcodes.py report.ncu-rep units-per-launch [top-N]
Prints executed warp-instructions x 32 / units per SASS opcode, plus stall samples and shared wavefronts per opcode."""
import csv
import subprocess
import sys
from collections import defaultdict

rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
inst, samp, wave = defaultdict(float), defaultdict(float), defaultdict(float)
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break  # first kernel only
        hdr = r
        ie, sm, src = r.index("Instructions Executed"), r.index("# Samples"), r.index("Source")
        wv = r.index("L1 Wavefronts Shared")
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    toks = r[src].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.rstrip(";")
    key = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]
    inst[key] += float(r[ie] or 0)
    samp[key] += float(r[sm] or 0)
    wave[key] += float(r[wv] or 0)
ti, ts = sum(inst.values()), sum(samp.values()) or 1.0
print(f"total {ti:.0f} warp-instructions = {ti * 32 / units:.1f} thread-instructions per unit; shared wavefronts "
      f"{sum(wave.values()) * 32 / units / 32:.1f} per 32 units")
for k, v in sorted(inst.items(), key=lambda x: -x[1])[:top]:
    print(f"{k:12s} {v * 32 / units:7.1f} /unit  {100 * v / ti:5.1f}% inst  {100 * samp[k] / ts:5.1f}% samples"
          + (f"  {wave[k] / (units / 32):6.1f} wavefronts/32 units" if wave[k] else ""))
