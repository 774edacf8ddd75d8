"""torchrun script: the frame every rank ends up with through the fused raster + peer exchange (and through the
NCCL all-gather path) must equal the single-GPU frame bit for bit.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gs = importlib.import_module("aframe-gaussian-splatting_b200")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    sc = gs.scenes
    w, h, n = 1000, 562, 60000
    rows = gs.synth_splats(n, 4242)
    frames = [sc.make_frame(sc.orbit_camera(w, h, s), sc.demo_object(), w, h) for s in (0, 11, 23, 37, 59, 80, 101)]
    ok = True
    with gs.SplatContext(local) as ref_ctx, gs.SplatContext(local) as ctx:
        ref_ctx.push_splats(rows)
        ref = [ref_ctx.render(f, fmt=gs.GS_FORMAT_RGBA8, bg=(0.1, 0.2, 0.3, 1.0)).copy() for f in frames]
        ctx.push_splats(rows)
        ctx.set_shard(rank, world)
        # ---- fused raster + peer exchange ----
        handles = [None] * world
        dist.all_gather_object(handles, ctx.peer_export(w * h * 4))
        ctx.peer_import(rank, world, handles)
        ctx.render(frames[0], fmt=gs.GS_FORMAT_RGBA8)  # sizes the instance buffers
        dist.barrier()
        outs = [ctx.pinned_array((h, w, 4), np.uint8) for _ in frames]
        tickets = []
        for i, f in enumerate(frames):
            outs[i][...] = 0
            p = ctx.make_params(f, bg=(0.1, 0.2, 0.3, 1.0), fmt=gs.GS_FORMAT_RGBA8, flags=gs.GS_RENDER_OUT_PEER)
            tickets.append(ctx.render_async(p, outs[i].ctypes.data))
            if i >= 2:
                ctx.wait(tickets[i - 2])
        for t in tickets[-2:]:
            ctx.wait(t)
        for i in range(len(frames)):
            same = np.array_equal(outs[i], ref[i])
            ok &= same
            if not same:
                d = np.abs(outs[i].astype(int) - ref[i].astype(int))
                print(f"[rank {rank}] peer frame {i} differs: max {d.max()} at {np.unravel_index(d.argmax(), d.shape)} n={int((d > 0).sum())}", flush=True)
        # ---- NCCL all-gather of tiles + un-tiling ----
        r, frame_t = gs.dist.make_gpu_sharded_renderer(ctx, frames[2], rank, world)
        out = r.render(frames[2])
        ctx.synchronize()
        got = out.cpu().numpy().reshape(h, w, 4)
        ref2 = ref_ctx.render(frames[2], fmt=gs.GS_FORMAT_RGBA8)
        same = np.array_equal(got, ref2)
        ok &= same
        if not same:
            print(f"[rank {rank}] nccl frame differs", flush=True)
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(flag.item()) == 1 else "MULTI_GPU_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
