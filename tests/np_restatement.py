"""Independent numpy restatement of the reference's pack / sort / projection, written separately from
oracle/gs_oracle.c (vectorised, different code shape) to cross-check the oracle, since the reference has
no tests or golden vectors (parity unpinned).  Follows index.js line by line like the oracle does."""
import numpy as np


def np_sort(matrices, view, cutout=None):
    """sortSplats (index.js:507-570) with numpy: fp64 arithmetic on f32 inputs, stable argsort of the keys."""
    m = np.asarray(matrices, np.float32).reshape(-1, 16).astype(np.float64)
    v = np.asarray(view, np.float32).astype(np.float64)
    x, y, z, s = m[:, 12], m[:, 13], m[:, 14], m[:, 15]
    depth = ((v[0] * x + v[1] * y) + v[2] * z) + v[3]
    keep = (depth < 0) & (s > -0.0001 * depth)
    if cutout is not None:
        e = np.asarray(cutout, np.float32).astype(np.float64)
        ny = -y
        with np.errstate(divide="ignore", invalid="ignore"):
            w = 1.0 / (((e[3] * x + e[7] * ny) + e[11] * z) + e[15])
            c0 = (((e[0] * x + e[4] * ny) + e[8] * z) + e[12]) * w
            c1 = (((e[1] * x + e[5] * ny) + e[9] * z) + e[13]) * w
            c2 = (((e[2] * x + e[6] * ny) + e[10] * z) + e[14]) * w
        out = (c0 < -0.5) | (c0 > 0.5) | (c1 < -0.5) | (c1 > 0.5) | (c2 < -0.5) | (c2 > 0.5)
        keep &= ~out
    idx = np.nonzero(keep)[0]
    d = depth[idx]
    if len(idx) == 0:
        return np.zeros(0, np.uint32)
    mn, mx = d.min(), d.max()
    d32 = d.astype(np.float32).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 65535.0 / (mx - mn)
        val = (d32 - mn) * inv
    val = np.where(np.isfinite(val), val, 0.0)
    # ToInt32: truncate, wrap modulo 2^32, reinterpret as signed (fmod of doubles is exact)
    w = np.fmod(np.trunc(val), 4294967296.0)
    w = np.where(w < 0, w + 4294967296.0, w)
    key = w.astype(np.int64)
    key = np.where(key >= 2147483648, key - 4294967296, key)
    ok = (key >= 0) & (key <= 65535)
    order = np.argsort(key[ok], kind="stable")
    out = np.zeros(len(idx), np.uint32)
    out[: ok.sum()] = idx[ok][order]
    return out


def np_pack(rows):
    """pushDataBuffer (index.js:343-402) vectorised; parseInt handled as trunc (valid when no |v| < 1e-6, v != 0)."""
    rows = np.asarray(rows, np.uint8).reshape(-1, 32)
    n = rows.shape[0]
    f = rows[:, :24].copy().view(np.float32).reshape(n, 6).astype(np.float64)
    b = rows[:, 28:32].astype(np.float64)
    qw, qx, qy, qz = (b[:, 0] - 128) / 128, (b[:, 1] - 128) / 128, (b[:, 2] - 128) / 128, -((b[:, 3] - 128) / 128)
    x2, y2, z2 = qx + qx, qy + qy, qz + qz
    xx, xy, xz = qx * x2, qx * y2, qx * z2
    yy, yz, zz = qy * y2, qy * z2, qz * z2
    wx, wy, wz = qw * x2, qw * y2, qw * z2
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - (yy + zz); R[:, 1, 0] = xy + wz; R[:, 2, 0] = xz - wy
    R[:, 0, 1] = xy - wz; R[:, 1, 1] = 1 - (xx + zz); R[:, 2, 1] = yz + wx
    R[:, 0, 2] = xz + wy; R[:, 1, 2] = yz - wx; R[:, 2, 2] = 1 - (xx + yy)
    s = f[:, 3:6]
    A = np.transpose(R, (0, 2, 1)) * s[:, None, :]  # R^T with column k scaled by s_k
    sig = np.empty((n, 3, 3))
    for r in range(3):
        for c in range(3):
            sig[:, r, c] = (A[:, r, 0] * A[:, c, 0] + A[:, r, 1] * A[:, c, 1]) + A[:, r, 2] * A[:, c, 2]
    six = np.stack([sig[:, 0, 0], sig[:, 1, 0], sig[:, 2, 0], sig[:, 1, 1], sig[:, 2, 1], sig[:, 2, 2]], axis=1)
    mx = np.abs(six).max(axis=1)
    cs = np.stack([f[:, 0], f[:, 1], -f[:, 2], mx / 32767.0], axis=1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = six * 32767.0 / mx[:, None]
    tiny = (np.abs(q) < 1e-6) & (q != 0)
    i16 = np.where(np.isfinite(q), np.trunc(q), 0).astype(np.int64).astype(np.int16)
    u16 = i16.view(np.uint16).astype(np.uint32)
    cc = np.empty((n, 4), np.uint32)
    cc[:, 0] = u16[:, 0] | (u16[:, 1] << 16)
    cc[:, 1] = u16[:, 2] | (u16[:, 3] << 16)
    cc[:, 2] = u16[:, 4] | (u16[:, 5] << 16)
    cc[:, 3] = rows[:, 24:28].copy().view(np.uint32).reshape(n)
    sa = (s.max(axis=1) * rows[:, 27].astype(np.float64) / 255.0).astype(np.float32)
    return cs, cc, sa, tiny.any(axis=1)


def np_project(cs, cc, proj, mv, width, height, focal):
    """Vertex shader (index.js:101-164) in float32 numpy with the same left-to-right sums.  Returns a dict of
    per-splat arrays (visible mask, cx, cy, v1, v2)."""
    f32 = np.float32
    cs = np.asarray(cs, f32).reshape(-1, 4)
    cc = np.asarray(cc, np.uint32).reshape(-1, 4)
    P = np.asarray(proj, f32)
    M = np.asarray(mv, f32)
    x, y, z = cs[:, 0], cs[:, 1], cs[:, 2]
    one = f32(1.0)
    cam = [((M[r] * x + M[4 + r] * y) + M[8 + r] * z) + M[12 + r] * one for r in range(4)]
    p = [((P[r] * cam[0] + P[4 + r] * cam[1]) + P[8 + r] * cam[2]) + P[12 + r] * cam[3] for r in range(4)]
    bounds = f32(1.2) * p[3]
    culled = (p[2] < -p[3]) | (p[0] < -bounds) | (p[0] > bounds) | (p[1] < -bounds) | (p[1] > bounds)

    def unpack(u):
        lo = (u & 0xFFFF).astype(np.int64)
        lo = np.where(lo >= 32768, lo - 65536, lo)
        hi = (u.astype(np.int64) << 32 >> 48)  # arithmetic shift of the signed 32-bit value
        return lo.astype(f32), hi.astype(f32)

    s = cs[:, 3]
    c00, c01 = unpack(cc[:, 0]); c02, c11 = unpack(cc[:, 1]); c12, c22 = unpack(cc[:, 2])
    c00, c01, c02, c11, c12, c22 = [c * s for c in (c00, c01, c02, c11, c12, c22)]
    V = [[c00, c01, c02], [c01, c11, c12], [c02, c12, c22]]
    fo = f32(focal)
    zero = np.zeros_like(x)
    with np.errstate(all="ignore"):
        zz = cam[2] * cam[2]
        J = [[fo / cam[2], zero, zero], [zero, -fo / cam[2], zero], [-(fo * cam[0]) / zz, (fo * cam[1]) / zz, zero]]  # J[row][col]
        W = [[M[r * 4 + c] for c in range(3)] for r in range(3)]
        T = [[(W[r][0] * J[0][c] + W[r][1] * J[1][c]) + W[r][2] * J[2][c] for c in range(3)] for r in range(3)]
        U = [[(T[0][r] * V[0][c] + T[1][r] * V[1][c]) + T[2][r] * V[2][c] for c in range(3)] for r in range(3)]
        cov00 = (U[0][0] * T[0][0] + U[0][1] * T[1][0]) + U[0][2] * T[2][0]
        cov10 = (U[1][0] * T[0][0] + U[1][1] * T[1][0]) + U[1][2] * T[2][0]
        cov11 = (U[1][0] * T[0][1] + U[1][1] * T[1][1]) + U[1][2] * T[2][1]
        vcx, vcy = p[0] / p[3], p[1] / p[3]
        d1 = cov00 + f32(0.3); off = cov10; d2 = cov11 + f32(0.3)
        mid = f32(0.5) * (d1 + d2)
        hd = (d1 - d2) / f32(2.0)
        radius = np.sqrt(hd * hd + off * off)
        l1 = mid + radius
        l2 = np.where((mid - radius) < f32(0.1), f32(0.1), mid - radius)
        dvx0, dvy0 = off, l1 - d1
        dlen = np.sqrt(dvx0 * dvx0 + dvy0 * dvy0)
        dvx, dvy = dvx0 / dlen, dvy0 / dlen
        s1, s2 = np.sqrt(f32(2.0) * l1), np.sqrt(f32(2.0) * l2)
        s1 = np.where(f32(1024.0) < s1, f32(1024.0), s1)
        s2 = np.where(f32(1024.0) < s2, f32(1024.0), s2)
        v1x, v1y = s1 * dvx, s1 * dvy
        v2x, v2y = s2 * dvy, s2 * (-dvx)
        zndc = p[2] / p[3]
        cx = (vcx * f32(0.5) + f32(0.5)) * f32(width)
        cy = (vcy * f32(0.5) + f32(0.5)) * f32(height)
    finite = np.isfinite(v1x) & np.isfinite(v1y) & np.isfinite(v2x) & np.isfinite(v2y) & np.isfinite(cx) & np.isfinite(cy)
    visible = (~culled) & (zndc <= 1.0) & finite
    return dict(visible=visible, cx=cx, cy=cy, v1x=v1x, v1y=v1y, v2x=v2x, v2y=v2y)
