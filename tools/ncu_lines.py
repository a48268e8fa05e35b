"""Join an `ncu --page source --csv` export with `nvdisasm -g` line info: executed instructions and stall samples
per source line.  usage: python tools/ncu_lines.py <source.csv> <cubin> <kernel-substring> [top]"""
import csv
import re
import subprocess
import sys


def sass_lines(cubin, kern):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    out, cur, on = [], ("?", 0), False
    for ln in txt:
        if ln.startswith(".text."):
            on = kern in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main():
    path, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    isamp, iinst = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    data = [(int(r[isamp]), int(r[iinst])) for r in rows[2:] if len(r) > iinst and r[iinst].isdigit()]
    sl = sass_lines(cubin, kern)
    if len(sl) != len(data):
        print("warning: %d SASS lines in the cubin vs %d in the report" % (len(sl), len(data)))
    agg = {}
    for (loc, _), (s, i) in zip(sl, data):
        a = agg.setdefault(loc, [0, 0, 0])
        a[0] += i
        a[1] += s
        a[2] += 1
    ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    src = {}
    print("total warp instructions %d, samples %d" % (ti, ts))
    for loc, (i, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if loc[0] not in src:
            try:
                src[loc[0]] = open("psalm_b200/csrc/" + loc[0]).read().splitlines()
            except OSError:
                src[loc[0]] = []
        text = src[loc[0]][loc[1] - 1].strip()[:90] if 0 < loc[1] <= len(src[loc[0]]) else ""
        print("%5.2f%% inst %5.2f%% samp %4d sass  %s:%d  %s" % (100.0 * i / ti, 100.0 * s / max(ts, 1), n, loc[0], loc[1], text))


if __name__ == "__main__":
    main()
