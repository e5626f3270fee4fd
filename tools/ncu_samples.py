"""Where a kernel WAITS: warp-state samples per CUDA source line, summed over every captured launch of the kernel.

usage: python tools/ncu_samples.py report.ncu-rep kernel-substring [top-N]
Needs -lineinfo and `ncu --set full --import-source on` (add `--warp-sampling-interval 0` for short kernels).
tools/ncu_lines.py is the companion that ranks lines by executed instructions.
"""
import csv
import subprocess
import sys


def main():
    rep, kfilter = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    lines, launches, grab, ie, sm = {}, 0, False, None, None
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "Function Name":
            grab = kfilter in r[1]
            launches += int(grab)
            continue
        if r[0] == "Line No":
            ie, sm = r.index("Instructions Executed"), r.index("# Samples")
            continue
        if grab and r[0].isdigit() and r[2] == "-":  # a CUDA source line (SASS rows carry an address)
            try:
                c, s = float(r[ie] or 0), float(r[sm] or 0)
            except ValueError:
                continue
            old = lines.get(int(r[0]), (0.0, 0.0, ""))
            lines[int(r[0])] = (old[0] + c, old[1] + s, r[1].strip())
    itot = sum(v[0] for v in lines.values()) or 1.0
    stot = sum(v[1] for v in lines.values()) or 1.0
    print(f"{kfilter}: {launches} launches, {stot:.0f} samples, {itot / max(launches, 1):.0f} warp-instructions per launch")
    for ln, (c, s, src) in sorted(lines.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{ln:5d}  samples {100 * s / stot:5.1f}%  inst {100 * c / itot:5.1f}%  {src[:120]}")


if __name__ == "__main__":
    main()
