"""Generates the golden fixtures of tests/golden/ from the oracle (oracle/gs_oracle.c).

The reference has no tests, golden vectors or runnable implementation in this image (parity unpinned), so
these vectors pin our own restatement: a hand-checkable 64-splat scene and a 20 k-splat seeded scene
(index arrays + 256x144 frames).  Run from the repo root:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
gs = importlib.import_module("aframe-gaussian-splatting_b200")
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
W, H = 256, 144
sc = gs.scenes


def main():
    # (1) 64 splats on a 4x4x4 lattice with distinct colours / scales / rotations
    n = 64
    rows = np.zeros((n, 32), np.uint8)
    k = np.arange(n)
    pos = np.stack([(k % 4) - 1.5, ((k // 4) % 4) * 0.6 - 0.9, (k // 16) * 0.8 - 1.2], axis=1).astype(np.float32)
    scl = np.stack([0.05 + 0.01 * (k % 5), 0.08 - 0.01 * (k % 3), 0.03 + 0.02 * (k % 2)], axis=1).astype(np.float32)
    rows[:, 0:12] = pos.view(np.uint8).reshape(n, 12)
    rows[:, 12:24] = scl.view(np.uint8).reshape(n, 12)
    rows[:, 24] = (k * 37) % 256; rows[:, 25] = (k * 91) % 256; rows[:, 26] = (k * 13) % 256; rows[:, 27] = 60 + (k * 3) % 196
    rows[:, 28] = 128 + (k * 7) % 120; rows[:, 29] = 128 - (k * 5) % 100; rows[:, 30] = 128 + (k * 11) % 90; rows[:, 31] = 128 - (k * 3) % 80
    cs, cc, m = orc.pack(rows)
    fr = sc.make_frame(sc.fixed_camera(W, H), sc.demo_object(), W, H, sc.demo_cutout())
    # a tighter cutout so that it actually removes splats of the lattice
    cut = gs.three_math.Object3D(position=sc.DEMO_OBJECT_POSITION, scale=(2.2, 1.5, 1.5))
    cutm = np.asarray(gs.three_math.world_to_cutout(cut, sc.demo_object()).elements, np.float32)
    order = orc.sort(m, fr.view)
    order_c = orc.sort(m, fr.view, cutm)
    img, _ = orc.render(cs, cc, order, fr.proj, fr.modelview, W, H, fr.focal)
    np.savez_compressed(os.path.join(HERE, "scene64.npz"), rows=rows, center_scale=cs, cov_color=cc, size_alpha=m[:, 15],
                        view=fr.view, proj=fr.proj, modelview=fr.modelview, cutout=cutm, order=order, order_cutout=order_c,
                        width=W, height=H, focal=np.float32(fr.focal), frame=img)
    print("scene64: V =", len(order), "V(cutout) =", len(order_c), "frame max", img.max())

    # (2) 20 k seeded synthetic scene (rows are regenerated from the seed; only results are stored)
    n2, seed = 20000, 0x5EED0001
    rows2 = gs.synth_splats(n2, seed)
    cs2, cc2, m2 = orc.pack(rows2)
    order2 = orc.sort(m2, fr.view)
    img2, st = orc.render(cs2, cc2, order2, fr.proj, fr.modelview, W, H, fr.focal)
    np.savez_compressed(os.path.join(HERE, "scene20k.npz"), n=n2, seed=seed, view=fr.view, proj=fr.proj, modelview=fr.modelview,
                        order=order2, width=W, height=H, focal=np.float32(fr.focal), frame=img2.astype(np.float16),
                        cov_xor=np.uint32(np.bitwise_xor.reduce(cc2.reshape(-1))))
    print("scene20k: V =", len(order2), st)


if __name__ == "__main__":
    main()
