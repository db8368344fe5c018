"""Model of the L1 data-stage cost of the march's corner-row loads on C3 (tools only, CPU).

The rule is the one MEASURED by tools/l1_gather_probe.cu on a B200 (profiles/r02_l1_gather_probe.txt):
  * a warp-wide load is served in passes of 128 bytes' worth of CONSECUTIVE lanes: 4 lanes per pass for LDG.256
    (8 passes), 8 lanes for LDG.128 (4 passes), 16 lanes for LDG.64;
  * lanes of a pass that read the same entry are a broadcast (free); different entries cost nothing extra as long as
    they sit in different 16/32-byte slots (bank groups) of their 128-byte lines -- whether or not the lines differ;
  * entries of DIFFERENT lines in the SAME slot serialise: a pass takes max over slots of the distinct entries there;
  * at most 4 distinct lines are looked up per cycle (only matters for passes of 8+ lanes).
For a layout (entries per line S, slot map) and a lane -> pixel arrangement inside the 8x4 warp tile the script
reports the mean data-stage cycles per warp-wide corner load over a sample of warps and steps of the C3 turntable.
"""
import itertools
import sys

import numpy as np

W, H, N = 1920, 1080, 512
STEP = 0.005           # world units per step (high_quality), bounds +-1 => 1.28 voxels


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


def t_near(pos, d):
    with np.errstate(divide="ignore"):
        inv = 1.0 / d
    a, b = (-1 - pos) * inv, (1 - pos) * inv
    return np.maximum(np.minimum(a, b).max(axis=1), 0.0)


L = np.arange(32)
ARR = {
    "rows 8x1 (4-lane pass = 4x1 pixels)": (L % 8, L // 8),
    "4x2 quarter blocks (pass = 4x1), round-2 default": (L % 4 + 4 * ((L // 8) % 2), (L // 4) % 2 + 2 * (L // 16)),
    "2x2 pass blocks, quarter = 4x2": (L % 2 + 2 * ((L // 4) % 2) + 4 * ((L // 8) % 2), (L // 2) % 2 + 2 * (L // 16)),
    "2x2 pass blocks, quarter = 2x4": (L % 2 + 2 * ((L // 8) % 4), (L // 2) % 2 + 2 * ((L // 4) % 2)),
}


def sample_warps(arr, n_tiles=40, views=(0, 20, 45, 70, 100, 135), jitter=0, seed=0):
    """Lower-tap cells (n_warp_loads, 32, 3) of warps over the object, lanes in lattice lock-step (+- jitter steps)."""
    rng = np.random.default_rng(seed)
    lx, ly = arr
    out = []
    for k in views:
        for _ in range(n_tiles):
            tx, ty = rng.integers(700, 1220), rng.integers(300, 780)
            px, py = (tx // 8) * 8 + lx, (ty // 4) * 4 + ly
            pos, d = rays(k, px.astype(float), py.astype(float))
            tn = t_near(pos, d)
            i0 = rng.integers(80, 400)
            for s in range(3):
                i = i0 + s + (rng.integers(-jitter, jitter + 1, 32) if jitter else 0)
                p = pos[None, :] + d * (tn + i * STEP)[:, None]
                out.append(np.floor((p + 1) / 2 * N - 0.5).astype(int))
    return np.array(out)


def cost(cells, lanes_per_pass, line_of, slot_of):
    """Mean cycles per warp-wide load."""
    total = 0.0
    for warp in cells:
        for g in range(0, 32, lanes_per_pass):
            uniq = {tuple(v) for v in warp[g:g + lanes_per_pass]}
            per_slot, lines = {}, set()
            for c in uniq:
                per_slot[slot_of(*c)] = per_slot.get(slot_of(*c), 0) + 1
                lines.add(line_of(*c))
            total += max(max(per_slot.values()), len(lines) / 4.0)
    return total / len(cells)


def layouts(S):
    """name -> (line_of, slot_of) for S entries per 128-byte line."""
    out = {}
    for rx, ry in itertools.product(range(S), repeat=2):       # pitched rows: slot = (iz + rx*ix + ry*iy) mod S
        out[f"pitched ({rx},{ry})"] = (
            (lambda x, y, z, rx=rx, ry=ry: (x, y, (z + rx * x + ry * y) // S)),
            (lambda x, y, z, rx=rx, ry=ry: (z + rx * x + ry * y) % S))
    if S == 4:      # 32-byte z-pair entries: a line is a 2x2 (x, y) block of entries of one z
        out["brick 2x2x1"] = ((lambda x, y, z: (x >> 1, y >> 1, z)), (lambda x, y, z: (x & 1) * 2 + (y & 1)))
        out["brick 2x1x2"] = ((lambda x, y, z: (x >> 1, y, z >> 1)), (lambda x, y, z: (x & 1) * 2 + (z & 1)))
        out["brick 1x2x2"] = ((lambda x, y, z: (x, y >> 1, z >> 1)), (lambda x, y, z: (y & 1) * 2 + (z & 1)))
        out["parity xor"] = ((lambda x, y, z: (x, y, z >> 2)), (lambda x, y, z: (z + 2 * (x & 1) + (y & 1) * 1 + 2 * (y & 1)) % 4))
    if S == 8:      # 16-byte entries: a line is a 2x2x2 block
        out["brick 2x2x2"] = ((lambda x, y, z: (x >> 1, y >> 1, z >> 1)), (lambda x, y, z: (x & 1) * 4 + (y & 1) * 2 + (z & 1)))
    return out


def main():
    quick = "--quick" in sys.argv
    for jitter in (0, 2):
        print(f"==== lanes {'in lattice lock-step' if jitter == 0 else f'+-{jitter} steps apart (after empty-space skipping)'}")
        for name, arr in ARR.items():
            cells = sample_warps(arr, n_tiles=10 if quick else 40, jitter=jitter)
            distinct = np.mean([len({tuple(v) for v in w}) for w in cells])
            print(f"-- {name}: {distinct:.1f} distinct entries per warp-load")
            for S, lanes, label in ((4, 4, "LDG.256, 32 B entries (f32x4 z-pair), floor 8"),
                                    (8, 8, "LDG.128, 16 B entries (f16x4 z-pair / f32x4), floor 4")):
                res = sorted((cost(cells, lanes, lo, so), n) for n, (lo, so) in layouts(S).items())
                pick = [r for r in res if r[1] in ("pitched (0,0)", "pitched (3,1)", "pitched (1,3)", "pitched (2,1)")
                        or r[1].startswith("brick") or r[1].startswith("parity")]
                print(f"   {label}")
                print("     best: " + ", ".join(f"{n} {c:.2f}" for c, n in res[:4]))
                print("     " + ", ".join(f"{n} {c:.2f}" for c, n in pick))


if __name__ == "__main__":
    main()
