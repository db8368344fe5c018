"""Second, independent restatement of pyvr/shaders/volume.frag.glsl in vectorised numpy (float32).

Used only to cross-check oracle/pyvr_oracle.c on small cases (tests/test_oracle.py): it shares no
code with the C oracle, evaluates every pixel in lock-step per loop iteration and uses plain
(unfused) float32 numpy arithmetic, so agreement is to ~1e-6, not bit-exact.
"""

import numpy as np

F = np.float32


def _taps(u, n):
    x = u * F(n) - F(0.5)
    fl = np.floor(x)
    f = (x - fl).astype(F)
    i = fl.astype(np.int64)
    return np.clip(i, 0, n - 1), np.clip(i + 1, 0, n - 1), f


def _tex3d(buf, u, v, s):
    """buf: (d, h, w[, c]) array = GL texture with width w fastest; coords u (width), v, s (depth)."""
    d, h, w = buf.shape[:3]
    i0, i1, fx = _taps(u, w)
    j0, j1, fy = _taps(v, h)
    k0, k1, fz = _taps(s, d)
    if buf.ndim == 4:
        fx, fy, fz = fx[:, None], fy[:, None], fz[:, None]
    lerp = lambda a, b, t: a + t * (b - a)  # noqa: E731
    x00 = lerp(buf[k0, j0, i0], buf[k0, j0, i1], fx)
    x10 = lerp(buf[k0, j1, i0], buf[k0, j1, i1], fx)
    x01 = lerp(buf[k1, j0, i0], buf[k1, j0, i1], fx)
    x11 = lerp(buf[k1, j1, i0], buf[k1, j1, i1], fx)
    return lerp(lerp(x00, x10, fy), lerp(x01, x11, fy), fz).astype(F)


def render(volume, camera, light, config, lut, width, height):
    """Returns (rgba8 (H,W,4) uint8 bottom-row-first, accum (H,W,4) float32, samples)."""
    data = np.ascontiguousarray(volume.data, dtype=F)
    # moderngl texture3d(shape): (w, h, d) = shape over C-order bytes -> numpy view (d, h, w)
    s0, s1, s2 = data.shape
    tex = data.reshape(-1).reshape(s2, s1, s0)
    ntex = None
    if volume.normals is not None:
        ntex = np.ascontiguousarray(volume.normals, dtype=F).reshape(-1).reshape(s2, s1, s0, 3)
    lut = np.ascontiguousarray(lut, dtype=F)
    size = lut.shape[0]

    # uniforms: matrix.tobytes() read column-major by GL -> mathematical matrix = numpy array transposed
    V = camera.get_view_matrix().astype(F).T
    P = camera.get_projection_matrix(width / height).astype(F).T
    iV, iP = np.linalg.inv(V).astype(F), np.linalg.inv(P).astype(F)
    pos = np.asarray(camera.get_camera_vectors()[0], dtype=F)

    py, px = np.mgrid[0:height, 0:width]
    uvx = ((px.ravel().astype(F) + F(0.5)) / F(width)).astype(F)
    uvy = ((py.ravel().astype(F) + F(0.5)) / F(height)).astype(F)
    clip = np.stack([uvx * F(2) - F(1), uvy * F(2) - F(1), -np.ones_like(uvx), np.ones_like(uvx)], axis=0)
    eye = (iP @ clip).astype(F)
    eye[2], eye[3] = F(-1), F(0)
    wdir = (iV @ eye)[:3].astype(F)
    wdir = (wdir / np.sqrt((wdir * wdir).sum(axis=0))).astype(F)          # (3, N)

    bmin = np.asarray(volume.min_bounds, dtype=F)[:, None]
    bmax = np.asarray(volume.max_bounds, dtype=F)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (F(1) / wdir).astype(F)
        tmin = ((bmin - pos[:, None]) * inv).astype(F)
        tmax = ((bmax - pos[:, None]) * inv).astype(F)
    t_near = np.fmax(np.fmax(np.fmin(tmin, tmax)[0], np.fmin(tmin, tmax)[1]), np.fmin(tmin, tmax)[2])
    t_far = np.fmin(np.fmin(np.fmax(tmin, tmax)[0], np.fmax(tmin, tmax)[1]), np.fmax(tmin, tmax)[2])
    hit = (t_near <= t_far) & (t_far > 0)
    t_near = np.maximum(t_near, F(0)).astype(F)

    n = uvx.size
    p = (pos[:, None] + wdir * t_near[None, :]).astype(F)
    acc = np.zeros((n, 4), dtype=F)
    acc_a = np.zeros(n, dtype=F)
    ldir = (np.asarray(light.target, dtype=F) - np.asarray(light.position, dtype=F)).astype(F)
    ldir = (ldir / np.sqrt((ldir * ldir).sum())).astype(F)
    step, ref = F(config.step_size), F(config.reference_step_size)
    dstep = (wdir * step).astype(F)
    samples = 0
    for _ in range(config.max_steps):
        live = hit & (acc_a < F(0.99))
        if not live.any():
            break
        tc = ((p - bmin) / (bmax - bmin)).astype(F)
        u, v, s = tc[2], tc[1], tc[0]                                     # x <-> z swizzle
        ok = live & (u >= 0) & (u <= 1) & (v >= 0) & (v <= 1) & (s >= 0) & (s <= 1)
        idx = np.nonzero(ok)[0]
        if idx.size:
            samples += idx.size
            dens = _tex3d(tex, u[idx], v[idx], s[idx])
            l0, l1, lf = _taps(dens, size)
            rgba = (lut[l0] + lf[:, None] * (lut[l1] - lut[l0])).astype(F)
            alpha = (F(1) - np.exp(-rgba[:, 3] * step / ref)).astype(F)
            if ntex is not None:
                nrm = _tex3d(ntex, u[idx], v[idx], s[idx])
            else:
                nrm = np.stack([dens, np.zeros_like(dens), np.zeros_like(dens)], axis=1)
            with np.errstate(divide="ignore", invalid="ignore"):
                nrm = (nrm / np.sqrt((nrm * nrm).sum(axis=1, keepdims=True))).astype(F)
            diff = np.fmax((nrm * ldir[None, :]).sum(axis=1), F(0)).astype(F)
            lightv = (F(light.ambient_intensity) + F(light.diffuse_intensity) * diff).astype(F)
            one_minus = (F(1) - acc_a[idx]).astype(F)
            acc[idx, :3] += one_minus[:, None] * (rgba[:, :3] * lightv[:, None] * alpha[:, None])
            acc_a[idx] += one_minus * alpha
        p = (p + dstep).astype(F)
    acc[:, 3] = acc_a
    a = np.clip(acc[:, 3], 0, 1)
    out = np.concatenate([np.clip(acc[:, :3], 0, 1) * a[:, None], (a * a)[:, None]], axis=1)
    rgba8 = np.rint(np.clip(out, 0, 1) * F(255)).astype(np.uint8)
    return rgba8.reshape(height, width, 4), acc.reshape(height, width, 4), samples
