#!/usr/bin/env python
"""Brute-force search of the linear L1 slot maps used by the packed texel layout (csrc/common.cuh).

slot(ix, iy, iz) = (a*ix + b*iy + c*iz) mod SLOTS, c odd so that the SLOTS z-consecutive texels of a line
fill it.  Two texels at offset d collide (same banks of different lines) iff a*dx + b*dy + c*dz = 0 mod SLOTS.
Ranks the triples by the length of the shortest collision vector, then by how few vectors have it.
"""
import sys

for mod in (4, 8, 16):
    rng = range(-4, 5)
    vecs = sorted(((x, y, z) for x in rng for y in rng for z in rng if (x, y, z) != (0, 0, 0)),
                  key=lambda v: v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    res = []
    for a in range(mod):
        for b in range(mod):
            for c in range(1, mod, 2):
                for v in vecs:
                    if (a * v[0] + b * v[1] + c * v[2]) % mod == 0:
                        n2 = v[0] ** 2 + v[1] ** 2 + v[2] ** 2
                        cnt = sum(1 for w in vecs if w[0] ** 2 + w[1] ** 2 + w[2] ** 2 == n2 and (a * w[0] + b * w[1] + c * w[2]) % mod == 0)
                        res.append((n2, -cnt, a, b, c))
                        break
    res.sort(reverse=True)
    print(f"SLOTS={mod}: best (|d|^2, -count, a, b, c) = {res[:4]}")
