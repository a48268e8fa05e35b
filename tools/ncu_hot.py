"""Top SASS instructions of an `ncu --page source --csv` export by stall samples and executed count."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, isamp, iinst = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    iconf = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[isamp]), int(r[iinst]), int(r[iconf]) if iconf is not None else 0, r[isrc].strip()))
        except (ValueError, IndexError):
            pass
    ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
    print("instructions (warp-level) %d, samples %d, SASS lines %d" % (ti, ts, len(data)))
    print("--- by stall samples")
    for s, i, c, src in sorted(data, reverse=True)[:top]:
        print("%6.2f%% samp %6.2f%% inst  excess_smem_wf %9d  %s" % (100.0 * s / max(ts, 1), 100.0 * i / max(ti, 1), c, src[:100]))
    ops = {}
    for s, i, c, src in data:
        op = src.split()[0] if src and not src.startswith("@") else (src.split()[1] if len(src.split()) > 1 else src)
        ops.setdefault(op.split(".")[0], [0, 0])
        ops[op.split(".")[0]][0] += i
        ops[op.split(".")[0]][1] += s
    print("--- by opcode")
    for op, (i, s) in sorted(ops.items(), key=lambda x: -x[1][0])[:18]:
        print("%-10s %6.2f%% inst %6.2f%% samp" % (op, 100.0 * i / max(ti, 1), 100.0 * s / max(ts, 1)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
