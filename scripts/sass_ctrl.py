#!/usr/bin/env python
"""Decode the scheduling control bits of a cuobjdump -sass listing (sm_70+ 128-bit encoding): per instruction the
stall count, yield, write / read scoreboard it SETS and the scoreboard mask it WAITS on.
  cuobjdump -sass -fun <mangled> lib.so | python scripts/sass_ctrl.py [regex]
Used to check that the look-ahead corner loads and the loads being consumed sit on different scoreboards."""
import re, sys
rx = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
lines = sys.stdin.read().splitlines()
i = 0
n = 0
while i < len(lines):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = (hi >> 41) & 0x1FFFFF
            stall = ctrl & 0xF
            yld = (ctrl >> 4) & 1
            wr = (ctrl >> 5) & 7
            rd = (ctrl >> 8) & 7
            wait = (ctrl >> 11) & 0x3F
            txt = m.group(2)
            if rx is None or rx.search(txt):
                w = ",".join(str(b) for b in range(6) if wait >> b & 1)
                print(f"{n:5d} {m.group(1)} st={stall:2d} {'Y' if yld else ' '} wr={'-' if wr == 7 else wr} rd={'-' if rd == 7 else rd} wait={{{w}}}  {txt}")
            n += 1
            i += 2
            continue
    i += 1
