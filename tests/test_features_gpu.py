"""GPU tests of the draw's optional features through the C ABI: depth interop (index.js:179-180), the two pixel loops
of the raster (packed fp32x2 / scalar) producing identical frames, and the GS_RENDER_STATS counters."""
import os

import numpy as np
import pytest

from conftest import scene_inputs

pytestmark = pytest.mark.gpu
FRAME_TOL = 1e-3


def _depth_plane(orc, cs, cc, order, fr, w, h):
    """A depth buffer that splits the scene: left half at the median splat depth, a ramp on the right, far plane in a
    corner block (so that all three regimes - everything rejected, partially rejected, nothing rejected - occur)."""
    p = orc.project(cs, cc, order, fr.proj, fr.modelview, w, h, fr.focal)
    zw = (p["zndc"][p["visible"] == 1] * np.float32(0.5) + np.float32(0.5)).astype(np.float32)
    lo, mid, hi = np.percentile(zw, [5, 50, 95]).astype(np.float32)
    d = np.empty((h, w), np.float32)
    d[:, : w // 2] = mid
    d[:, w // 2:] = np.linspace(lo, hi, w - w // 2, dtype=np.float32)[None, :]
    d[: h // 4, : w // 4] = 1.0
    d[-h // 4:, -w // 4:] = 0.0
    return d


@pytest.mark.parametrize("n,w,h", [(20000, 256, 144), (120000, 1000, 562)])
def test_depth_interop_parity(gs, orc, ctx, n, w, h):
    rows, cs, cc, m, fr = scene_inputs(gs, orc, n, 900 + n, w, h)
    ctx.clear(); ctx.push_packed(cs, cc, m[:, 15])
    order = orc.sort(m, fr.view)
    depth = _depth_plane(orc, cs, cc, order, fr, w, h)
    bg = (0.3, 0.2, 0.1, 1.0)
    exp, est = orc.render(cs, cc, order, fr.proj, fr.modelview, w, h, fr.focal, bg=bg, depth_in=depth)
    base, bst = orc.render(cs, cc, order, fr.proj, fr.modelview, w, h, fr.focal, bg=bg)
    assert 0 < est["fragments"] < bst["fragments"]
    got = ctx.render(fr, bg=bg, fmt=gs.GS_FORMAT_RGBA32F, depth_in=depth)
    err = np.abs(got - exp)
    assert err.max() <= FRAME_TOL, (float(err.max()), np.unravel_index(err.argmax(), err.shape))
    # the block at depth 0 shows only the background, the block at depth 1 equals the frame without a depth buffer
    assert np.allclose(got[-h // 4:, -w // 4:], bg, atol=1e-6)
    plain = ctx.render(fr, bg=bg, fmt=gs.GS_FORMAT_RGBA32F)
    assert np.array_equal(got[: h // 4, : w // 4], plain[: h // 4, : w // 4])
    # device-resident depth buffer (GS_RENDER_DEPTH_DEVICE) gives the same frame
    import torch
    t = torch.from_numpy(depth).cuda()
    torch.cuda.synchronize()
    p = ctx.make_params(fr, bg=bg, fmt=gs.GS_FORMAT_RGBA32F, flags=gs.GS_RENDER_DEPTH_DEVICE)
    p.depth_in = t.data_ptr()
    out = np.empty((h, w, 4), np.float32)
    ctx.render_raw(p, out.ctypes.data)
    assert np.array_equal(out, got)
    # RGBA8 output of the depth-tested frame
    got8 = ctx.render(fr, bg=bg, fmt=gs.GS_FORMAT_RGBA8, depth_in=depth)
    e8 = np.floor(np.clip(exp, 0, 1) * 255.0 + 0.5).astype(np.int32)
    assert np.abs(got8.astype(np.int32) - e8).max() <= 2


def test_packed_and_scalar_pixel_loops_agree(gs, orc):
    """GS_RASTER=scalar selects the one-pixel-per-lane loop; the default is the packed fp32x2 loop.  Same operations in
    the same order per pixel -> bit-identical frames (float and RGBA8, with and without a depth buffer)."""
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 150000, 4321, 1000, 562)
    order = orc.sort(m, fr.view)
    depth = _depth_plane(orc, cs, cc, order, fr, 1000, 562)
    frames = {}
    old = os.environ.get("GS_RASTER")
    try:
        for mode in ("scalar", "packed"):
            os.environ["GS_RASTER"] = mode
            with gs.SplatContext(0) as c:
                c.push_packed(cs, cc, m[:, 15])
                frames[mode] = (c.render(fr, fmt=gs.GS_FORMAT_RGBA32F, bg=(0.1, 0.2, 0.3, 0.4)).copy(),
                                c.render(fr, fmt=gs.GS_FORMAT_RGBA8).copy(),
                                c.render(fr, fmt=gs.GS_FORMAT_RGBA32F, depth_in=depth).copy())
    finally:
        if old is None:
            os.environ.pop("GS_RASTER", None)
        else:
            os.environ["GS_RASTER"] = old
    for a, b in zip(frames["scalar"], frames["packed"]):
        assert np.array_equal(a, b)
    exp, _ = orc.render(cs, cc, order, fr.proj, fr.modelview, 1000, 562, fr.focal, bg=(0.1, 0.2, 0.3, 0.4))
    assert np.abs(frames["packed"][0] - exp).max() <= FRAME_TOL


def test_stats_frame_counts(gs, orc, ctx):
    """GS_RENDER_STATS: D = number of (splat, 16x16 tile) pairs whose tile meets the r<=2 footprint, and the pair
    counters.  Checked against counts derived from the oracle's projection on the host."""
    w, h = 640, 360
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 30000, 77, w, h)
    ctx.clear(); ctx.push_packed(cs, cc, m[:, 15])
    order = orc.sort(m, fr.view)
    plain = ctx.render(fr, fmt=gs.GS_FORMAT_RGBA32F).copy()
    st0 = ctx.stats()
    assert st0["n_tile_instances"] == 0 and st0["n_pair_tests"] == 0  # only filled by a STATS frame
    got = ctx.render(fr, fmt=gs.GS_FORMAT_RGBA32F, stats=True)
    st = ctx.stats()
    assert np.array_equal(got, plain)  # the statistics frame renders the same picture
    # host count: tiles (of the visible splats in the draw order) containing at least one covered pixel centre
    p = orc.project(cs, cc, order, fr.proj, fr.modelview, w, h, fr.focal)
    exact = 0
    pairs = 0
    for s in p[p["visible"] == 1]:
        ex = 2 * np.hypot(s["v1x"], s["v2x"]) + 1; ey = 2 * np.hypot(s["v1y"], s["v2y"]) + 1
        x0 = max(0, int(np.floor(s["cx"] - ex))); x1 = min(w - 1, int(np.ceil(s["cx"] + ex)))
        y0 = max(0, int(np.floor(s["cy"] - ey))); y1 = min(h - 1, int(np.ceil(s["cy"] + ey)))
        if x0 > x1 or y0 > y1:
            continue
        dx = (np.arange(x0, x1 + 1, dtype=np.float32) + np.float32(0.5)) - s["cx"]
        dy = (np.arange(y0, y1 + 1, dtype=np.float32) + np.float32(0.5)) - s["cy"]
        DX, DY = np.meshgrid(dx, dy)
        px = DX * s["a2x"] + DY * s["a2y"]; py = DX * s["a1x"] + DY * s["a1y"]
        msk = (px * px + py * py) <= 4.0
        if not msk.any():
            continue
        yy, xx = np.nonzero(msk)
        exact += len(np.unique(((yy + y0) >> 4) * 4096 + ((xx + x0) >> 4)))
        pairs += int(msk.sum())
    # the cull is conservative (0.5 % slack, closest point on the tile box): it may keep a few tiles no pixel centre covers
    assert exact <= st["n_tile_instances"] <= exact * 1.05 + 16
    assert st["n_instances_kept"] < st["n_tile_instances"]            # 64x64 bins: fewer instances than tiles
    assert st["n_records_streamed"] >= st["n_tile_instances"]
    assert 0 < st["n_pair_hits"] <= st["n_pair_tests"]
    assert st["n_pair_hits"] <= pairs * 1.001 + 16                     # early-stopped pixels skip pairs, never add any


def _eye_cameras(gs, w, h, ipd=0.064):
    """Two eye cameras around the fixed head camera (same orientation, +-ipd/2 along x)."""
    tm = gs.three_math
    head = gs.scenes.fixed_camera(w, h)
    eyes = []
    for sx in (-0.5, 0.5):
        eyes.append(tm.PerspectiveCamera(fov=80.0, aspect=w / h, near=0.005, far=10000.0,
                                         position=(head.position[0] + sx * ipd, head.position[1], head.position[2])))
    return head, eyes


def test_stereo_one_sort_two_eyes(gs, orc, ctx):
    """gs_render_stereo (index.js:184-195 per-eye onBeforeRender + one tick() sort): both eyes are drawn with the HEAD
    camera's order; each eye frame equals the oracle's frame for (head order, eye matrices)."""
    w, h = 640, 400
    rows, cs, cc, m, fr_head = scene_inputs(gs, orc, 60000, 2024, w, h)
    ctx.clear(); ctx.push_packed(cs, cc, m[:, 15])
    sc = gs.scenes
    head, eye_cams = _eye_cameras(gs, w, h)
    fr_head = sc.make_frame(head, sc.demo_object(), w, h)
    eyes = [sc.make_frame(c, sc.demo_object(), w, h) for c in eye_cams]
    order = orc.sort(m, fr_head.view)
    got = ctx.render_stereo(fr_head.view, eyes, fmt=gs.GS_FORMAT_RGBA32F, bg=(0.0, 0.1, 0.2, 1.0))
    for e, g in zip(eyes, got):
        exp, _ = orc.render(cs, cc, order, e.proj, e.modelview, w, h, e.focal, bg=(0.0, 0.1, 0.2, 1.0))
        assert np.abs(g - exp).max() <= FRAME_TOL
    assert not np.array_equal(got[0], got[1])
    st = ctx.last_stereo_stats
    assert st[0].n_sorted == st[1].n_sorted == len(order) and st[0].ms_sort > 0 and st[1].ms_sort == 0
    # with a cutout the cull acts in the one sort
    cut_fr = sc.make_frame(head, sc.demo_object(), w, h, sc.demo_cutout())
    order_c = orc.sort(m, cut_fr.view, cut_fr.cutout)
    got = ctx.render_stereo(cut_fr.view, eyes, cutout=cut_fr.cutout, fmt=gs.GS_FORMAT_RGBA32F)
    exp, _ = orc.render(cs, cc, order_c, eyes[1].proj, eyes[1].modelview, w, h, eyes[1].focal)
    assert len(order_c) < len(order) and np.abs(got[1] - exp).max() <= FRAME_TOL


def test_component_render_xr(gs, orc):
    """The component mirror: xrPixelRatio scales the eye viewports (index.js:13-15), tick()'s camera sorts."""
    sc = gs.scenes
    rows = gs.synth_splats(30000, 11)
    cs, cc, m = orc.pack(rows)
    w, h = 800, 450
    head, eye_cams = _eye_cameras(gs, w, h)
    comp = gs.GaussianSplattingComponent({"src": rows.tobytes(), "xrPixelRatio": 0.5})
    comp.init(head, sc.demo_object())
    try:
        left, right = comp.render_xr(eye_cams, w, h, fmt=gs.GS_FORMAT_RGBA32F)
        assert left.shape == (225, 400, 4) and right.shape == (225, 400, 4)
        fr_head = sc.make_frame(head, sc.demo_object(), 400, 225)
        order = orc.sort(m, fr_head.view)
        e = sc.make_frame(eye_cams[0], sc.demo_object(), 400, 225)
        exp, _ = orc.render(cs, cc, order, e.proj, e.modelview, 400, 225, e.focal)
        assert np.abs(left - exp).max() <= FRAME_TOL
    finally:
        comp.renderer.close()


def test_progressive_push_while_rendering(gs, orc):
    """index.js:259-298: rows are pushed as they arrive while the scene is drawn.  Pushes are interleaved with
    gs_render_async; every frame must equal the oracle's frame of the prefix that was resident when it was submitted."""
    w, h = 640, 360
    n, chunk = 240000, 40000
    rows = gs.synth_splats(n, 555)
    cs, cc, m = orc.pack(rows)
    sc = gs.scenes
    fr = sc.make_frame(sc.fixed_camera(w, h), sc.demo_object(), w, h)
    with gs.SplatContext(0) as c:
        c.reserve(n)  # initGL(numVertexes): no growth (hence no pipeline wait) during the load
        outs, tickets, prefixes = [], [], []
        for first in range(0, n, chunk):
            c.push_splats(rows[first:first + chunk])
            out = c.pinned_array((h, w, 4), np.float32)
            out[...] = -1.0
            t = c.render_async(c.make_params(fr, fmt=gs.GS_FORMAT_RGBA32F), out.ctypes.data)
            outs.append(out); tickets.append(t); prefixes.append(first + chunk)
            if len(tickets) >= 3:  # keep three frames in flight across the pushes
                st = c.wait(tickets[-3])
                assert st.n_splats == prefixes[-3]
        for t, k in zip(tickets[-2:], prefixes[-2:]):
            assert c.wait(t).n_splats == k
        for out, k in zip(outs, prefixes):
            order = orc.sort(m[:k], fr.view)
            exp, _ = orc.render(cs[:k], cc[:k], order, fr.proj, fr.modelview, w, h, fr.focal)
            assert np.abs(out - exp).max() <= FRAME_TOL, k
        assert c.num_splats == n
        # without gs_reserve the table grows geometrically; frames stay correct across the growth
    with gs.SplatContext(0) as c:
        outs, tickets, prefixes = [], [], []
        for first in range(0, n, chunk):
            c.push_splats(rows[first:first + chunk])
            out = np.empty((h, w, 4), np.float32)
            tickets.append(c.render_async(c.make_params(fr, fmt=gs.GS_FORMAT_RGBA32F), out.ctypes.data))
            outs.append(out); prefixes.append(first + chunk)
        for t in tickets[-3:]:
            c.wait(t)
        for out, k in zip(outs[::2], prefixes[::2]):
            order = orc.sort(m[:k], fr.view)
            exp, _ = orc.render(cs[:k], cc[:k], order, fr.proj, fr.modelview, w, h, fr.focal)
            assert np.abs(out - exp).max() <= FRAME_TOL, k
