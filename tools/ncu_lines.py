#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample totals from an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py <rep> <kernel regex> [top N] [--by-samples]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    fname = None
    out = []
    tot_inst = tot_samp = 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) >= 8 and r[0].isdigit():
            try:
                inst = int(r[7])
                samp = int(r[4])
                tinst = int(r[8])
            except ValueError:
                continue
            out.append((inst, samp, fname, int(r[0]), r[1].strip()[:100], tinst))
            tot_inst += inst
            tot_samp += samp
    print("total warp instructions %d, stall samples %d" % (tot_inst, tot_samp))
    if "--by-samples" in sys.argv:
        out.sort(key=lambda t: t[1], reverse=True)
    else:
        out.sort(reverse=True)
    for inst, samp, f, ln, src, tinst in out[:top]:
        print("%6.2f%% inst %6.2f%% samp %5.1f lanes  %s:%d  %s" % (100.0 * inst / max(tot_inst, 1), 100.0 * samp / max(tot_samp, 1),
                                                                    tinst / max(inst, 1), f, ln, src))


if __name__ == "__main__":
    main()
