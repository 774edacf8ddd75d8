"""GPU probe: where does the e2e time go? (device-only vs host-copy pipelines, with/without L2 flush)"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
gs = importlib.import_module("aframe-gaussian-splatting_b200")
sc = gs.scenes
n, w, h, seed, _ = sc.CONFIGS["train_1m_1080p"]
rows = gs.synth_splats(n, seed)
fr = sc.make_frame(sc.fixed_camera(w, h), sc.demo_object(), w, h)
ctx = gs.SplatContext(0)
ctx.push_splats(rows)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx._lib.gs_stream(ctx._h), device=dev)
with torch.cuda.stream(stream):
    fdev = [torch.zeros(h * w * 4, dtype=torch.uint8, device=dev) for _ in range(3)]
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
host = [ctx.pinned_array((h, w, 4), np.uint8) for _ in range(3)]
pd = ctx.make_params(fr, flags=gs.GS_RENDER_OUT_DEVICE)
ph = ctx.make_params(fr, flags=0)

def run(kind, steps=40, do_flush=False, depth=2):
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tickets = []
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        r0.record(stream)
    for i in range(steps):
        if do_flush:
            with torch.cuda.stream(stream):
                flush.zero_()
        if kind == "dev":
            tickets.append(ctx.render_async(pd, fdev[i % 3].data_ptr()))
        else:
            tickets.append(ctx.render_async(ph, host[i % 3].ctypes.data))
        if i >= depth - 1:
            ctx.wait(tickets[i - (depth - 1)])
    for t in tickets[-(depth - 1):] if depth > 1 else []:
        ctx.wait(t)
    with torch.cuda.stream(stream):
        r1.record(stream)
    stream.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    return r0.elapsed_time(r1) / steps, wall / steps

for _ in range(2):
    for kind in ("dev", "host"):
        for fl in (False, True):
            for depth in (1, 2, 3):
                ev, wall = run(kind, do_flush=fl, depth=depth)
                print(f"{kind:5s} flush={fl!s:5s} depth={depth}: {ev:.4f} ms/frame (events)  {wall:.4f} ms/frame (wall)  dev ms_total={ctx.last_stats.ms_total:.4f}")
# raw D2H bandwidth
with torch.cuda.stream(stream):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hp = torch.from_numpy(host[0].reshape(-1))
    e0.record(stream)
    for _ in range(10):
        hp.copy_(fdev[0], non_blocking=True)
    e1.record(stream)
stream.synchronize()
print("D2H 8.3 MB:", e0.elapsed_time(e1) / 10, "ms ->", h * w * 4 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9, "GB/s")
