"""Summarise an `ncu --page source --csv` SASS dump: share of executed instructions and stall
samples per block of SASS lines, with the dominant opcodes (to locate hot loops)."""
import collections
import csv
import sys


def main(path, block=250):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    ix_src, ix_inst, ix_samp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows:
        if len(r) >= len(hdr) and r[ix_inst].isdigit():
            data.append((r[ix_src].strip(), int(r[ix_inst]), int(r[ix_samp])))
    tot = sum(d[1] for d in data)
    ts = sum(d[2] for d in data) or 1
    print("total warp-inst", tot, "samples", ts, "sass lines", len(data))
    for b in range(0, len(data), block):
        blk = data[b:b + block]
        inst = sum(d[1] for d in blk)
        samp = sum(d[2] for d in blk)
        ops = collections.Counter()
        for s, i, _ in blk:
            tok = s.split()
            op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
            ops[op.split(".")[0]] += i
        print(f"{b:6d} inst {inst / tot:6.3f} samp {samp / ts:6.3f} ",
              ", ".join(f"{k}:{v / max(inst, 1):.2f}" for k, v in ops.most_common(7)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 250)
