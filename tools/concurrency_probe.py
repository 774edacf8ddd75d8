"""How much GPU time does one context's three-stage pipeline leave unused?  Drive K independent SplatContexts (same
scene, each with three frames in flight) from one host thread and compare the aggregate frame rate with K = 1.
    python tools/concurrency_probe.py [K ...]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gs = importlib.import_module("aframe-gaussian-splatting_b200")


def main():
    ks = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
    sc = gs.scenes
    n, w, h, seed, _ = sc.CONFIGS["train_1m_1080p"]
    rows = gs.synth_splats(n, seed)
    fr = sc.make_frame(sc.fixed_camera(w, h), sc.demo_object(), w, h)
    for K in ks:
        ctxs = [gs.SplatContext(0) for _ in range(K)]
        bufs = []
        for c in ctxs:
            c.push_splats(rows)
            bufs.append([torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda") for _ in range(3)])
        torch.cuda.synchronize()
        ps = [c.make_params(fr, fmt=gs.GS_FORMAT_RGBA8, flags=gs.GS_RENDER_OUT_DEVICE) for c in ctxs]

        def run(steps):
            tickets = [[] for _ in ctxs]
            for i in range(steps):
                for j, c in enumerate(ctxs):
                    tickets[j].append(c.render_async(ps[j], bufs[j][i % 3].data_ptr()))
                    if i >= 2:
                        c.wait(tickets[j][i - 2])
            for j, c in enumerate(ctxs):
                for t in tickets[j][-2:]:
                    c.wait(t)
        run(10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 200
        run(steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"K={K}: {K * steps / dt:8.1f} frames/s aggregate ({1e3 * dt / steps:.3f} ms per round of {K} frames)", flush=True)
        for c in ctxs:
            c.close()


if __name__ == "__main__":
    main()
