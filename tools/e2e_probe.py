"""Where does the end-to-end frame time go?  Config 2, device-resident vs host frames, several pipeline depths, with and
without the L2 flush; prints frames/s (host clock around K frames + synchronize) and the mean per-stage times."""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gs = importlib.import_module("aframe-gaussian-splatting_b200")


def main():
    sc = gs.scenes
    n, w, h, seed, cutout = sc.CONFIGS["train_1m_1080p"]
    rows = gs.synth_splats(n, seed)
    fr = sc.make_frame(sc.fixed_camera(w, h), sc.demo_object(), w, h)
    ctx = gs.SplatContext(0)
    ctx.reserve(n)
    ctx.push_splats(rows)
    dev = torch.device("cuda:0")
    stream = torch.cuda.ExternalStream(ctx._lib.gs_stream(ctx._h), device=dev)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    p_dev = ctx.make_params(fr, fmt=gs.GS_FORMAT_RGBA8, flags=gs.GS_RENDER_OUT_DEVICE)
    p_host = ctx.make_params(fr, fmt=gs.GS_FORMAT_RGBA8, flags=0)
    frames_dev = [torch.zeros(h * w * 4, dtype=torch.uint8, device=dev) for _ in range(4)]
    host_gs = [ctx.pinned_array((h, w, 4), np.uint8) for _ in range(4)]
    host_torch = [torch.empty(h * w * 4, dtype=torch.uint8).pin_memory() for _ in range(4)]

    def run(k, depth, host, do_flush, torch_pinned=False):
        tickets, stats = [], []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(k):
            if do_flush:
                with torch.cuda.stream(stream):
                    flush.zero_()
            if host:
                dst = host_torch[i % 4].data_ptr() if torch_pinned else host_gs[i % 4].ctypes.data
                tickets.append(ctx.render_async(p_host, dst))
            else:
                tickets.append(ctx.render_async(p_dev, frames_dev[i % 4].data_ptr()))
            if i >= depth - 1:
                stats.append(ctx.wait(tickets[i - (depth - 1)]).as_dict())
        for t in tickets[max(0, len(tickets) - (depth - 1)):]:
            stats.append(ctx.wait(t).as_dict())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        m = {q: float(np.mean([s[q] for s in stats[k // 4:]])) for q in ("ms_sort", "ms_project", "ms_bin", "ms_raster", "ms_total")}
        return k / dt, m

    run(400, 3, False, True)  # warm
    for name, kw in (("device d3 flush", dict(depth=3, host=False, do_flush=True)),
                     ("device d3 noflush", dict(depth=3, host=False, do_flush=False)),
                     ("host d2 flush", dict(depth=2, host=True, do_flush=True)),
                     ("host d3 flush", dict(depth=3, host=True, do_flush=True)),
                     ("host d4 flush", dict(depth=4, host=True, do_flush=True)),
                     ("host d4 noflush", dict(depth=4, host=True, do_flush=False)),
                     ("host d4 flush torch-pinned", dict(depth=4, host=True, do_flush=True, torch_pinned=True)),
                     ("device d3 flush", dict(depth=3, host=False, do_flush=True))):
        best = None
        for _ in range(3):
            fps, m = run(400, **kw)
            if best is None or fps > best[0]:
                best = (fps, m)
        print(f"{name:28s} {best[0]:8.1f} frames/s  " + " ".join(f"{q[3:]} {v:.4f}" for q, v in best[1].items()), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
