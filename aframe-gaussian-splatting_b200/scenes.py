"""Deterministic synthetic scenes and cameras.  The reference's datasets
(`train.splat`, `bicycle.ply`) are named only by URL (index.html:13, README.md:47) and are absent, so
every configuration is synthetic: 32-byte `.splat` rows in the reference's on-wire layout
(index.js:671-676) produced by a counter-based RNG (splitmix64), so any language can regenerate them.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

from .three_math import Object3D, PerspectiveCamera, Matrix4, get_model_view_matrix, get_projection_matrix, \
    world_to_cutout, focal_length, yaw_quaternion

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser of a uint64 counter array."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _uniform(seed: int, stream: int, n: int, start: int = 0) -> np.ndarray:
    """n doubles in (0,1): counter = seed*2^40 + stream*2^32.. + index (53 random bits)."""
    with np.errstate(over="ignore"):
        base = np.uint64((seed * 0x1000003 + stream * 0x9E3779B1) & 0xFFFFFFFFFFFFFFFF)
        ctr = (np.arange(start, start + n, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D) + base) & _M64
    bits = splitmix64(ctr) >> np.uint64(11)
    return (bits.astype(np.float64) + 0.5) / float(1 << 53)


def _normal(seed: int, stream: int, n: int, start: int = 0) -> np.ndarray:
    u1 = _uniform(seed, stream, n, start)
    u2 = _uniform(seed, stream + 1000, n, start)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2)


BOX_LO = np.array([-4.0, -1.5, -4.0])
BOX_HI = np.array([4.0, 2.5, 4.0])
_N_SURF = 64
_BLOCK = 1 << 20  # rows per generation block (the RNG is counter-based, so blocks are independent)


def _surfaces(seed: int):
    sc = BOX_LO + _uniform(seed, 1, _N_SURF * 3).reshape(_N_SURF, 3) * (BOX_HI - BOX_LO)
    frame = _normal(seed, 2, _N_SURF * 9).reshape(_N_SURF, 3, 3)
    q, _ = np.linalg.qr(frame)
    ext = 0.3 + 1.2 * _uniform(seed, 4, _N_SURF * 3).reshape(_N_SURF, 3)
    ellipsoid = (np.arange(_N_SURF) % 2) == 1
    return sc, q, ext, ellipsoid


def _synth_block(args):
    """rows [start, start+n) of the scene in GENERATION order + their importance (f32)."""
    seed, start, n, log_scale_mean = args
    sc, q, ext, ellipsoid = _surfaces(seed)
    n_surf = _N_SURF
    sel = _uniform(seed, 5, n, start)
    sid = np.minimum((_uniform(seed, 6, n, start) * n_surf).astype(np.int64), n_surf - 1)
    a = _uniform(seed, 7, n, start) * 2.0 - 1.0
    b = _uniform(seed, 8, n, start) * 2.0 - 1.0
    # planar patch: centre + a*ext0*t0 + b*ext1*t1 (+ thin jitter along the normal)
    jitter = (_uniform(seed, 9, n, start) - 0.5) * 0.02
    local = np.stack([a * ext[sid, 0], b * ext[sid, 1], jitter], axis=1)
    # ellipsoid shell: unit-sphere direction scaled by the radii
    theta = 2.0 * math.pi * _uniform(seed, 10, n, start)
    cz = _uniform(seed, 11, n, start) * 2.0 - 1.0
    sr = np.sqrt(np.maximum(0.0, 1.0 - cz * cz))
    sph = np.stack([sr * np.cos(theta) * ext[sid, 0], sr * np.sin(theta) * ext[sid, 1], cz * ext[sid, 2]], axis=1)
    local = np.where(ellipsoid[sid][:, None], sph, local)
    pos = sc[sid] + np.einsum("nij,nj->ni", q[sid], local)
    uni = BOX_LO + np.stack([_uniform(seed, 12, n, start), _uniform(seed, 13, n, start), _uniform(seed, 14, n, start)], axis=1) * (BOX_HI - BOX_LO)
    pos = np.where((sel < 0.8)[:, None], pos, uni)
    pos = np.clip(pos, BOX_LO, BOX_HI)

    scale = np.exp(log_scale_mean + 0.9 * np.stack([_normal(seed, 20, n, start), _normal(seed, 22, n, start), _normal(seed, 24, n, start)], axis=1))
    scale = np.clip(scale, 1e-4, 0.5)
    alpha = np.rint(255.0 / (1.0 + np.exp(-(0.5 + 2.5 * _normal(seed, 30, n, start))))).astype(np.uint8)
    rgb = np.minimum((np.stack([_uniform(seed, 40, n, start), _uniform(seed, 41, n, start), _uniform(seed, 42, n, start)], axis=1) * 256.0), 255.0).astype(np.uint8)
    quat = np.stack([_normal(seed, 50, n, start), _normal(seed, 52, n, start), _normal(seed, 54, n, start), _normal(seed, 56, n, start)], axis=1)
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    rot = np.clip(np.rint(quat * 128.0 + 128.0), 0, 255).astype(np.uint8)  # stored w,x,y,z

    rows = np.zeros((n, 32), dtype=np.uint8)
    rows[:, 0:12] = pos.astype(np.float32).view(np.uint8).reshape(n, 12)
    rows[:, 12:24] = scale.astype(np.float32).view(np.uint8).reshape(n, 12)
    rows[:, 24:27] = rgb
    rows[:, 27] = alpha
    rows[:, 28:32] = rot
    sf = scale.astype(np.float32).astype(np.float64)
    importance = (sf[:, 0] * sf[:, 1] * sf[:, 2] * (alpha.astype(np.float64) / 255.0)).astype(np.float32)
    return rows, importance


def synth_splats(n: int, seed: int, log_scale_mean: float | None = None, sort_by_importance: bool = True,
                 workers: int | None = None) -> np.ndarray:
    """Return (n, 32) uint8 `.splat` rows.

    positions: 80 % on 64 random planar / ellipsoidal surfaces inside the box, 20 % uniform in the box;
    per-axis scale exp(N(mu, 0.9^2)) clamped to [1e-4, 0.5] (mu = -4.6, or -5.3 from 20 M splats up);
    alpha byte round(255*sigmoid(N(0.5, 2.5^2))); rgb bytes uniform; rotation = normalised N(0,1)^4 as
    clamp(round(q*128+128), 0, 255) stored w,x,y,z; rows ordered by descending sx*sy*sz*alpha, the order
    `processPlyBuffer` gives real scenes (index.js:655-668).

    Rows are generated in independent blocks of 2^20 (counter-based RNG), on a process pool when the scene is
    large; the result does not depend on the block size or the worker count.
    """
    if log_scale_mean is None:
        log_scale_mean = -5.3 if n >= 20_000_000 else -4.6
    jobs = [(seed, s, min(_BLOCK, n - s), log_scale_mean) for s in range(0, n, _BLOCK)]
    if workers is None:
        workers = min(len(jobs), max(1, (os.cpu_count() or 1) // 2), 32)
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_synth_block, jobs, chunksize=1)
    else:
        parts = [_synth_block(j) for j in jobs]
    rows = np.concatenate([p[0] for p in parts]) if len(parts) > 1 else parts[0][0]
    if sort_by_importance:
        importance = np.concatenate([p[1] for p in parts]) if len(parts) > 1 else parts[0][1]
        order = np.argsort(-importance, kind="stable")
        rows = rows[order]
    return np.ascontiguousarray(rows)


@dataclass
class FrameInputs:
    """Everything one frame hands the hot path: the reference's per-frame uniforms (index.js:184-195) and
    the sort request (index.js:441-453), as f32 column-major arrays."""
    proj: np.ndarray         # (16,) f32  gsProjectionMatrix
    modelview: np.ndarray    # (16,) f32  gsModelViewMatrix
    view: np.ndarray         # (4,)  f32  row 2 of modelview
    width: int
    height: int
    focal: float
    cutout: np.ndarray | None = None  # (16,) f32 worldToCutout, or None


def make_frame(camera: PerspectiveCamera, obj: Object3D, width: int, height: int, cutout: Object3D | None = None) -> FrameInputs:
    """What `onBeforeRender` + `tick` compute for one draw (index.js:184-195, 438-453)."""
    proj = get_projection_matrix(camera)
    mv = get_model_view_matrix(camera, obj)
    mvf = np.asarray(mv.elements, dtype=np.float32)  # uniform upload rounds to f32
    view = np.array([mv.elements[2], mv.elements[6], mv.elements[10], mv.elements[14]], dtype=np.float32)  # index.js:442
    cut = None
    if cutout is not None:
        cut = np.asarray(world_to_cutout(cutout, obj).elements, dtype=np.float32)  # index.js:452
    return FrameInputs(proj=np.asarray(proj.elements, dtype=np.float32), modelview=mvf, view=view, width=width,
                       height=height, focal=float(np.float32(focal_length(height, proj))), cutout=cut)


# the demo's entity transform (index.html:13: position="0 1.5 -2")
DEMO_OBJECT_POSITION = (0.0, 1.5, -2.0)
# cutout-demo.html:23 box scale, centred on the scene
CUTOUT_SCALE = (4.17, 2.95, 3.89)


def demo_object() -> Object3D:
    return Object3D(position=DEMO_OBJECT_POSITION)


def fixed_camera(width: int, height: int) -> PerspectiveCamera:
    """Configs 1/2/4/5: A-Frame default camera at (0, 1.6, 0), identity rotation."""
    return PerspectiveCamera(fov=80.0, aspect=width / height, near=0.005, far=10000.0, position=(0.0, 1.6, 0.0))


def orbit_camera(width: int, height: int, step: int, steps: int = 120, radius: float = 3.0) -> PerspectiveCamera:
    """Config 3: 360 degree yaw orbit around the entity, looking at it."""
    theta = 2.0 * math.pi * step / steps
    cx, cz = DEMO_OBJECT_POSITION[0], DEMO_OBJECT_POSITION[2]
    pos = (cx + radius * math.sin(theta), 1.6, cz + radius * math.cos(theta))
    return PerspectiveCamera(fov=80.0, aspect=width / height, near=0.005, far=10000.0, position=pos,
                             quaternion=yaw_quaternion(theta))


def demo_cutout() -> Object3D:
    return Object3D(position=DEMO_OBJECT_POSITION, scale=CUTOUT_SCALE)


CONFIGS = {
    # name: (n_splats, width, height, seed, cutout)
    "train_1m_1080p": (1_000_000, 1920, 1080, 0x5EED0002, False),
    "bicycle_6m_1080p_orbit": (6_000_000, 1920, 1080, 0x5EED0003, False),
    "synth_20m_2160p_cutout": (20_000_000, 3840, 2160, 0x5EED0004, True),
    "synth_80m_1080p": (80_000_000, 1920, 1080, 0x5EED0005, False),
}
