"""Per-CUDA-source-line executed-instruction counts from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [pixels-per-launch]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kfilter = sys.argv[2] if len(sys.argv) > 2 else ""
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    fn, agg, grab = None, {}, False
    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            fn = r[1]
            grab = (kfilter in fn) and fn not in agg
            if grab:
                agg[fn] = {}
            continue
        if r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            sm = hdr.index("# Samples")
            continue
        if grab and r[0].isdigit() and r[2] == "-":  # a CUDA source line (SASS rows carry an address)
            try:
                agg[fn][int(r[0])] = (float(r[ie] or 0), float(r[sm] or 0), r[1].strip())
            except ValueError:
                pass
    for fn, lines in agg.items():
        tot = sum(v[0] for v in lines.values())
        stot = sum(v[1] for v in lines.values()) or 1.0
        print(f"== {fn[:110]}\n   warp-instructions {tot:.0f}" + (f" = {tot * 32 / units:.1f} per unit" if units else ""))
        for ln, (c, s, src) in sorted(lines.items(), key=lambda x: -x[1][0])[:40]:
            per = f"{c * 32 / units:7.1f}" if units else f"{100 * c / tot:6.1f}%"
            print(f"{ln:5d} {per} stall {100 * s / stot:5.1f}%  {src[:100]}")
        break


if __name__ == "__main__":
    main()
