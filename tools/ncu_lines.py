"""Executed warp instructions per CUDA source line of an ncu report captured with --import-source on:
    python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur = fpath = hdr = None
    agg, sass_tot = {}, {}
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur = r[1][:70]
        elif r[0] == "Line No":
            hdr = r
            ia = hdr.index("Instructions Executed")
        elif hdr and cur and r[0].isdigit():
            agg.setdefault(cur, {}).setdefault((fpath, int(r[0]), r[1].strip()[:100]), 0)
            agg[cur][(fpath, int(r[0]), r[1].strip()[:100])] += int(r[ia]) if len(r) > ia and r[ia].isdigit() else 0
        elif hdr and cur and r[0] == "" and len(r) > ia and r[ia].isdigit():
            sass_tot[cur] = sass_tot.get(cur, 0) + int(r[ia])
    for k, v in agg.items():
        tot = sass_tot.get(k, 1)
        print("=====", k, "warp instructions", tot)
        for (f, ln, src), n in sorted(v.items(), key=lambda x: -x[1])[:top]:
            print("  %5.1f%%  %s:%d  %s" % (100.0 * n / tot, f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
