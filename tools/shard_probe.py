"""GPU probe: per-stage device times of rank 0 of a 1/2/8-way tile-column sharding, emulated on one GPU."""
import importlib, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gs = importlib.import_module("aframe-gaussian-splatting_b200")
sc = gs.scenes
n, w, h, seed, _ = sc.CONFIGS["train_1m_1080p"]
rows = gs.synth_splats(n, seed)
fr = sc.make_frame(sc.fixed_camera(w, h), sc.demo_object(), w, h)
ctx = gs.SplatContext(0)
ctx.push_splats(rows)
for world in (1, 2, 8):
    ctx.set_shard(0, world)
    tiles = ctx.owned_tiles(w, h, 0, world)
    buf = ctx.device_alloc(tiles * 1024)
    p = ctx.make_params(fr, flags=gs.GS_RENDER_OUT_DEVICE | gs.GS_RENDER_OUT_TILED)
    for _ in range(4):
        st = ctx.render_raw(p, buf)
    print(world, {k: round(getattr(st, k), 4) for k in ("ms_sort", "ms_project", "ms_bin", "ms_raster", "ms_total")}, st.n_instances, st.n_instances_kept, flush=True)
    ctx.device_free(buf)
