"""Per-source-line hot spots from an ncu report (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur_file, hdr, items = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) >= len(hdr) - 2 and r[2] == "-" and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            try:
                ie = float(d["Instructions Executed"]); smp = float(d["# Samples"])
            except (KeyError, ValueError):
                continue
            stalls = {k[6:]: float(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "-")}
            items.append((ie, smp, cur_file, r[0], r[1].strip(), stalls, d))
    tot_i = sum(i[0] for i in items); tot_s = sum(i[1] for i in items)
    print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
    items.sort(key=lambda x: -x[1])
    print("  %smp  %inst  file:line  [top stalls]  source")
    for ie, smp, f, ln, src, stalls, d in items[:top]:
        ts = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        tss = " ".join(f"{k}={v / max(smp, 1) * 100:.0f}%" for k, v in ts if v > 0)
        print(f"{smp / tot_s * 100:6.2f} {ie / tot_i * 100:6.2f}  {f}:{ln:>4s}  [{tss}]  {src[:90]}")


if __name__ == "__main__":
    main()
