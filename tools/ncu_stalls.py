"""Warp-stall samples of one kernel per CUDA source line (needs a report taken with --import-source on and -lineinfo):
tools/ncu_stalls.py <rep> [top] > out.txt"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    if "Source" not in txt:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    kernel = rows[0][1] if rows and rows[0][0] == "Kernel Name" else "?"
    header = next(r for r in rows if "Warp Stall Sampling (All Samples)" in r)
    body = rows[rows.index(header) + 1:]
    isamp = header.index("Warp Stall Sampling (All Samples)")
    isrc = header.index("Source")
    per = collections.Counter()
    for r in body:
        # (the listing interleaves CUDA lines with their SASS: SASS rows have an address, CUDA rows a line number, in column 0)
        if len(r) > isamp and r[isamp].isdigit() and r[isrc].strip() and not r[0].startswith("0x"):
            per[r[isrc].strip()] += int(r[isamp])
    total = sum(per.values())
    print(f"{kernel}: warp stall samples per line of the listing ({total} samples)")
    for src, n in per.most_common(top):
        print(f"{n:8d} {100.0 * n / max(total, 1):5.1f}%  {src[:150]}")


if __name__ == "__main__":
    main()
