"""Summarise a clock64 timeline written by a -DSTTS_TC_TRACE_BUILD library (STTS_TC_TRACE=1|2|3, see conv_tc.cuh):
per promotion unit of each epilogue set the cycles spent waiting for the MMAs, promoting, waiting for the correction
accumulator, in the epilogue, and between units; per (tile, K-chunk) of the MMA issuer the wait / issue split.

usage: python tools/trace_summary.py trace.txt [block]      (block: which TRACE dump of the file, default the last)"""
import re
import sys


def parse(block):
    roles = {}
    for line in block.split("\n")[1:]:
        m = re.match(r"\s*(\w+):(.*)", line)
        if m:
            roles[m.group(1)] = {int(a): int(c) for a, c in re.findall(r"(\d+):(-?\d+)", m.group(2))}
    return roles


def main():
    txt = open(sys.argv[1]).read()
    blocks = txt.split("TRACE ")[1:]
    if not blocks:
        sys.exit("no TRACE block in " + sys.argv[1])
    b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else -1]
    print("TRACE", b.split("\n")[0])
    roles = parse(b)
    for name in ("set0", "set1"):
        d = roles.get(name, {})
        if not d:
            continue
        print(name, "per unit: wait_mma promote wait_corr epilogue gap | end")
        tot = [0, 0, 0, 0, 0]
        n = 0
        for q in sorted(set(i // 5 for i in d)):
            g = [d.get(q * 5 + k) for k in range(5)]
            nxt = d.get((q + 1) * 5)
            if None in g[:3]:
                continue
            row = [g[1] - g[0], g[2] - g[1], (g[3] - g[2]) if g[3] else 0, (g[4] - g[3]) if g[3] and g[4] else 0,
                   (nxt - g[4]) if nxt and g[4] else 0]
            print("  %3d: %6d %6d %6d %6d %6d | %s" % (q, *row, g[4]))
            if q > 0:
                tot = [a + b2 for a, b2 in zip(tot, row)]
                n += 1
        if n:
            print("  mean (units 1..): " + " ".join("%6d" % (t // n) for t in tot))
    d = roles.get("mma", {})
    if d:
        print("mma per (tile, K-chunk): wait_A wait_tmem issue")
        for g in sorted(set(i // 4 for i in d))[:24]:
            v = [d.get(g * 4 + k) for k in range(4)]
            if None in v:
                continue
            print("  %3d: %6d %6d %6d" % (g, v[1] - v[0], v[2] - v[1], v[3] - v[2]))


if __name__ == "__main__":
    main()
