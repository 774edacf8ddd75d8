"""Render a few frames of a bench workload between cudaProfilerStart / cudaProfilerStop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/profile_frames.py train_1m_1080p --frames 2
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_raster -o prof \
        python tools/profile_frames.py train_1m_1080p --frames 1

Frames are rendered one at a time (gs_render), so a launch list shows every kernel of a frame in order."""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gs = importlib.import_module("aframe-gaussian-splatting_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="train_1m_1080p")
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--splats", type=int, default=0)
    ap.add_argument("--stats", action="store_true", help="render GS_RENDER_STATS frames")
    ap.add_argument("--shard", default="", help="rank/world: render only this rank's share of the frame (tile-sharded mode on one GPU)")
    args = ap.parse_args()
    sc = gs.scenes
    n, w, h, seed, cutout = sc.CONFIGS[args.workload]
    if args.splats:
        n = args.splats
    rows = gs.synth_splats(n, seed)
    if "orbit" in args.workload:
        frames = [sc.make_frame(sc.orbit_camera(w, h, i), sc.demo_object(), w, h) for i in range(0, 120, 7)]
    else:
        frames = [sc.make_frame(sc.fixed_camera(w, h), sc.demo_object(), w, h, sc.demo_cutout() if cutout else None)]
    ctx = gs.SplatContext(0)
    if args.shard:
        r, wd = (int(v) for v in args.shard.split("/"))
        ctx.set_shard(r, wd)
    ctx.reserve(n)
    for first in range(0, n, 4 << 20):
        ctx.push_splats(rows[first:first + (4 << 20)])
    out = torch.zeros(h * w * 4, dtype=torch.uint8, device="cuda")
    flags = gs.GS_RENDER_OUT_DEVICE | (gs.GS_RENDER_STATS if args.stats else 0)
    ps = [ctx.make_params(f, fmt=gs.GS_FORMAT_RGBA8, flags=flags) for f in frames]
    for i in range(args.warm):
        ctx.render_raw(ps[i % len(ps)], out.data_ptr())
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(args.frames):
        st = ctx.render_raw(ps[i % len(ps)], out.data_ptr())
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print({k: getattr(st, k) for k in ("n_splats", "n_sorted", "n_visible", "n_instances", "n_instances_kept", "kernel_launches",
                                        "ms_sort", "ms_bin", "ms_raster", "ms_total")})
    ctx.close()


if __name__ == "__main__":
    main()
