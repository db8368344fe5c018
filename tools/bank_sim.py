"""Model of the L1 data-stage cost of the march's corner-row loads on C3 (tools only, CPU).

Assumption (matches the ncu counters of round 2: ~15 data-pipe wavefronts per LDG.256): a warp-wide load is
served quarter-warp by quarter-warp (lanes 8q..8q+7), 16 bytes per lane per pass, and a pass takes as many cycles
as the largest number of DISTINCT entries that fall on the same 16-byte bank group.  For a layout with S entries
per 128-byte line and a slot map slot(ix,iy,iz) the script reports, over a sample of warps and steps of the C3
turntable, the mean pass cost for several lane->pixel arrangements inside the 8x4 (or other) warp tile.
"""
import itertools
import sys

import numpy as np

W, H, N = 1920, 1080, 512
STEP = 0.005 * N / 2.0   # voxels per step (bounds +-1)


def camera(k):
    az, el, d = 2 * np.pi * k / 360, np.pi / 6, 3.0
    pos = d * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    t = np.tan(np.pi / 8)
    return pos, fwd, right * t * W / H, up * t


def rays(k, px, py):
    pos, fwd, r, u = camera(k)
    ndx = (px + 0.5) / W * 2 - 1
    ndy = (py + 0.5) / H * 2 - 1
    d = fwd[None, :] + ndx[:, None] * r[None, :] + ndy[:, None] * u[None, :]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return pos, d


ARR = {
    "8x1 rows (current)": lambda l: (l % 8, l // 8),
    "4x2 blocks": lambda l: (l % 4 + 4 * ((l // 8) % 2), (l // 4) % 2 + 2 * (l // 16)),
    "2x4 blocks": lambda l: (l % 2 + 2 * (l // 8), (l // 2) % 4),
}


def cost(entries, slots_of, S):
    """entries: (n_pass, 8, 3) int lower taps of the lanes of each pass; returns mean cycles per pass."""
    total = 0
    for e in entries:
        uniq = {tuple(v) for v in e}
        cnt = np.zeros(S, int)
        for (x, y, z) in uniq:
            cnt[slots_of(x, y, z) % S] += 1
        total += cnt.max()
    return total / len(entries)


def main():
    rng = np.random.default_rng(0)
    views = [0, 20, 45, 70, 100, 135]
    results = {}
    for name, arr in ARR.items():
        lane = np.arange(32)
        lx, ly = arr(lane)
        passes = []
        for k in views:
            for _ in range(60):
                tx, ty = rng.integers(700, 1220), rng.integers(300, 780)   # tiles over the object
                px, py = (tx // 8) * 8 + lx, (ty // 4) * 4 + ly
                pos, d = rays(k, px.astype(float), py.astype(float))
                t = 3.0 + rng.uniform(-0.4, 0.4)
                for s in range(3):
                    p = pos[None, :] + d * (t + s * 0.005)
                    vox = (p + 1) / 2 * N - 0.5
                    lo = np.floor(vox).astype(int)
                    for q in range(4):
                        passes.append(lo[8 * q:8 * q + 8])
        passes = np.array(passes)
        distinct = np.mean([len({tuple(v) for v in e}) for e in passes])
        row = {"distinct entries per quarter-warp": round(float(distinct), 2)}
        for S, label in ((4, "f32 z-pair (4 slots)"), (8, "f32 (8 slots)")):
            best = []
            for rx, ry in itertools.product(range(S), repeat=2):
                c = cost(passes, lambda x, y, z: rx * x + ry * y + z, S)
                best.append((c, rx, ry))
            best.sort()
            row[label] = {"best": best[:3], "none (0,0)": [b for b in best if b[1:] == (0, 0)][0][0],
                          "(1,3)": [b for b in best if b[1:] == (1, 3)][0][0], "(3,1)": [b for b in best if b[1:] == (3, 1)][0][0]}
        results[name] = row
        print(name, row, flush=True)


if __name__ == "__main__":
    main()
