#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).

    python tools/ncu_source_lines.py gpurun_out/x.ncu-rep k_tile_query [top]
"""
import csv
import io
import subprocess
import sys


def main(path, kernel, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file = None
    hdr = None
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "":
            continue  # SASS rows
        d = dict(zip(hdr, r))
        try:
            inst = int(d["Instructions Executed"])
            samp = int(d["# Samples"])
            thr = int(d["Thread Instructions Executed"])
        except (KeyError, ValueError):
            continue
        if inst or samp:
            lines.append((inst, samp, thr, cur_file, r[0], r[1].strip()[:110]))
    tot = sum(l[0] for l in lines) or 1
    tots = sum(l[1] for l in lines) or 1
    print(f"# {kernel}: {tot} warp instructions, {tots} samples; per source line (inlined code is attributed to the line it came from)")
    print(f"{'inst%':>6} {'samp%':>6} {'thr/inst':>8}  file:line  source")
    for inst, samp, thr, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * inst / tot:6.2f} {100 * samp / tots:6.2f} {thr / max(1, inst):8.1f}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
