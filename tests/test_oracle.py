"""CPU tests of the oracle (oracle/gs_oracle.c): cross-check against the independent numpy restatement
(tests/np_restatement.py), the committed golden vectors, and the edge cases Q1-Q12 of SURVEY.md A.6.
The reference ships no tests (package.json:7): these vectors pin OUR restatement, not the reference."""
import os

import numpy as np
import pytest

from conftest import scene_inputs
from np_restatement import np_pack, np_project, np_sort

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_pack_matches_numpy(gs, orc):
    rows = gs.synth_splats(20000, 11)
    cs, cc, m = orc.pack(rows)
    ncs, ncc, nsa, tiny = np_pack(rows)
    assert np.array_equal(cs.view(np.uint32), ncs.view(np.uint32))
    ok = ~tiny  # rows hitting the parseInt exponent-form quirk are checked separately
    assert ok.mean() > 0.99
    assert np.array_equal(cc[ok], ncc[ok])
    assert np.array_equal(m[:, 15].view(np.uint32), nsa.view(np.uint32))
    assert np.array_equal(m[:, 12:15], cs[:, :3])


def test_pack_parseint_quirk(orc):
    """Q2: parseInt(1.2e-7) == 1, parseInt(-9.5e-7) == -9, parseInt(NaN) -> 0, truncation toward zero (Q3)."""
    # identity rotation (w=255 -> 127/128, not exactly 1: Q1 no normalisation), anisotropic scale
    row = np.zeros((1, 32), np.uint8)
    row[0, 0:24] = np.array([1, 2, 3, 0.5, 0.25, 0.125], np.float32).view(np.uint8)
    row[0, 24:28] = [10, 20, 30, 200]
    row[0, 28:32] = [255, 128, 128, 128]
    cs, cc, m = orc.pack(row)
    w = 127 / 128
    r = 1.0  # xx = yy = zz = 0 -> R = I exactly even though |q| != 1
    assert cs[0, 2] == -3.0 and cs[0, 3] == np.float32(0.25 / 32767.0)
    lo = lambda u: np.int16(np.uint16(u & 0xFFFF))
    hi = lambda u: np.int16(np.uint16(u >> 16))
    assert lo(cc[0, 0]) == 32767 and hi(cc[0, 0]) == 0
    assert hi(cc[0, 1]) == int(0.0625 * 32767 / 0.25)  # trunc(8191.75) = 8191
    assert hi(cc[0, 2]) == int(0.015625 * 32767 / 0.25)
    assert m[0, 15] == np.float32(0.5 * 200 / 255.0)
    # all scales zero -> max_value 0 -> NaN -> stored 0
    row[0, 12:24] = 0
    cs, cc, m = orc.pack(row)
    assert cs[0, 3] == 0 and cc[0, 0] == 0 and cc[0, 1] == 0 and cc[0, 2] == 0


def test_pack_quirk_rows_leading_digit(gs, orc):
    """Rows whose scaled covariance entry falls in (0, 1e-6) must store its leading decimal digit."""
    rows = gs.synth_splats(200000, 12)
    cs, cc, m = orc.pack(rows)
    ncs, ncc, nsa, tiny = np_pack(rows)
    diff = np.any(cc != ncc, axis=1)
    assert not np.any(diff & ~tiny)  # differences only where the quirk applies
    # wherever they differ the oracle stored a digit 1..9 (or -1..-9) where trunc gives 0
    a = cc[diff][:, :3].copy().view(np.int16).reshape(-1, 6)
    b = ncc[diff][:, :3].copy().view(np.int16).reshape(-1, 6)
    ch = a != b
    assert np.all(b[ch] == 0) and np.all((np.abs(a[ch]) >= 1) & (np.abs(a[ch]) <= 9))


@pytest.mark.parametrize("cutout", [False, True])
def test_sort_matches_numpy(gs, orc, cutout):
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 50000, 21, 640, 360, cutout=cutout)
    a = orc.sort(m, fr.view, fr.cutout)
    b = np_sort(m, fr.view, fr.cutout)
    assert a.dtype == np.uint32 and np.array_equal(a, b)
    assert np.array_equal(orc.sort_compact(cs, m[:, 15], fr.view, fr.cutout), a)
    assert 0 < len(a) < len(rows)


def test_sort_is_back_to_front_and_stable(gs, orc):
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 30000, 22, 640, 360)
    o = orc.sort(m, fr.view)
    v = fr.view.astype(np.float64)
    depth = ((v[0] * m[o, 12].astype(np.float64) + v[1] * m[o, 13]) + v[2] * m[o, 14]) + v[3]
    mn, mx = depth.min(), depth.max()
    key = np.trunc((depth.astype(np.float32).astype(np.float64) - mn) * (65535.0 / (mx - mn))).astype(np.int64)
    assert np.all(np.diff(key) >= 0)  # farthest (most negative) first
    same = np.diff(key) == 0
    assert np.all(np.diff(o.astype(np.int64))[same] > 0)  # ties keep index order


def test_sort_edge_cases(orc):
    view = np.array([0, 0, 1, 0], np.float32)
    def mk(z, size=1.0):
        m = np.zeros((len(z), 16), np.float32)
        m[:, 14] = z
        m[:, 15] = size
        return m
    assert len(orc.sort(mk([]), view)) == 0                              # no splats
    assert len(orc.sort(mk([1.0, 2.0]), view)) == 0                      # Q6: V = 0 (all behind)
    assert np.array_equal(orc.sort(mk([-1.0]), view), [0])               # Q6: V = 1 -> depthInv inf, key NaN -> 0
    assert np.array_equal(orc.sort(mk([-2.0, -2.0, -2.0]), view), [0, 1, 2])  # all equal
    assert np.array_equal(orc.sort(mk([-1.0, -3.0, -2.0]), view), [1, 2, 0])  # back to front
    # size filter: size > 0.0001*|depth| strictly
    m = mk([-1.0, -1.0, -1.0]); m[:, 15] = [0.0001, 0.00010001, 0.0]
    assert np.array_equal(orc.sort(m, view), [0, 1]) or np.array_equal(orc.sort(m, view), [1])
    thr = -0.0001 * -1.0
    exp = [i for i in range(3) if float(np.float32(m[i, 15])) > thr]
    assert np.array_equal(orc.sort(m, view), exp)
    # Q12: cutout box, y negated, faces inclusive
    cut = np.eye(4, dtype=np.float32).T.reshape(16).copy()
    m = mk([-1.0, -1.0, -1.0, -1.0]); m[:, 14] = -0.25
    m[:, 12] = [0.5, 0.5000001, -0.5, 0.0]; m[:, 13] = [0.0, 0.0, 0.0, 0.6]
    assert np.array_equal(orc.sort(m, np.array([0, 0, 1, -1], np.float32), cut), [0, 2])


def test_sort_q5_key_out_of_range(orc):
    """Q5: f32 rounding of the depth can push a key past 65535 (or below 0); the reference silently drops the
    write, keeps length validCount and leaves the tail 0."""
    view = np.array([0, 0, 1, 0], np.float32)
    n = 64
    m = np.zeros((n, 16), np.float32)
    m[:, 15] = 1.0
    m[:, 14] = -1000.0 - np.arange(n, dtype=np.float64) * 1e-5  # range << f32 ulp at 1000 (6e-5)
    o = orc.sort(m, view)
    b = np_sort(m, view)
    assert len(o) == n and np.array_equal(o, b)
    # with view[3] != 0 the fp64 depth is not f32-representable -> keys scatter around and outside the range
    view2 = np.array([0, 0, 1, 1e-4], np.float32)
    o2, b2 = orc.sort(m, view2), np_sort(m, view2)
    assert np.array_equal(o2, b2)


def test_project_matches_numpy(gs, orc):
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 40000, 23, 960, 540)
    pr = orc.project(cs, cc, None, fr.proj, fr.modelview, fr.width, fr.height, fr.focal)
    ref = np_project(cs, cc, fr.proj, fr.modelview, fr.width, fr.height, fr.focal)
    assert np.array_equal(pr["visible"].astype(bool), ref["visible"])
    v = ref["visible"]
    assert 0.1 < v.mean() < 0.9
    for k in ("cx", "cy", "v1x", "v1y", "v2x", "v2y"):
        assert np.array_equal(pr[k][v].view(np.uint32), ref[k][v].view(np.uint32)), k
    # basis is orthogonal and positively oriented: det[v2 v1] > 0
    det = pr["v2x"][v] * pr["v1y"][v] - pr["v2y"][v] * pr["v1x"][v]
    assert np.all(det > 0)
    assert np.all(np.hypot(pr["v2x"][v], pr["v2y"][v]) >= np.float32(np.sqrt(0.2)) * 0.999)  # lambda2 floor 0.1
    assert np.all(np.hypot(pr["v1x"][v], pr["v1y"][v]) <= 1024.0 * 1.001)  # 1024 px cap


def test_render_single_splat_analytic(orc):
    """One axis-aligned splat at the screen centre: alpha = exp(-(dx^2/(2 sx2) + dy^2/(2 sy2))) * a with
    s*2 = cov + 0.3.  (sy2 > sx2 on purpose: with sx2 >= sy2 and no off-diagonal the reference's
    normalize(vec2(0,0)) is NaN and the splat vanishes - quirk Q8, checked below.)"""
    W, H = 64, 64
    scale = np.float32(0.01 / 32767.0)
    cs = np.array([[0, 0, -5.0, scale]], np.float32)
    q = lambda v: np.uint32(np.int16(v).view(np.uint16))
    s00 = 20970  # Sigma00 = 20970 * scale ~ 0.0064, Sigma11 = Sigma22 = 0.01
    cc = np.zeros((1, 4), np.uint32)
    cc[0, 0] = q(s00)                 # (Sigma00, Sigma01 = 0)
    cc[0, 1] = q(32767) << 16         # (Sigma02 = 0, Sigma11)
    cc[0, 2] = q(32767) << 16         # (Sigma12 = 0, Sigma22)
    cc[0, 3] = 255 | (128 << 8) | (0 << 16) | (255 << 24)  # r=255, g=128, b=0, a=255
    P = np.zeros(16, np.float32); P[0] = 1.0; P[5] = -1.0; P[10] = -1.0; P[11] = -1.0; P[14] = -0.02
    MV = np.eye(4, dtype=np.float32).reshape(16)
    focal = 200.0
    img, st = orc.render(cs, cc, np.array([0], np.uint32), P, MV, W, H, focal, nthreads=2)
    assert st["n_visible"] == 1
    k = (focal / 5.0) ** 2
    sx2 = float(np.float32(s00) * scale) * k + 0.3
    sy2 = float(np.float32(32767) * scale) * k + 0.3
    ys, xs = np.mgrid[0:H, 0:W]
    r2 = (xs + 0.5 - 32.0) ** 2 / (2 * sx2) + (ys + 0.5 - 32.0) ** 2 / (2 * sy2)
    exp_a = np.where(r2 <= 4.0, np.exp(-r2), 0.0)
    safe = np.abs(r2 - 4.0) > 0.05  # pixels on the r2 == 4 contour may flip under f32 rounding
    assert (exp_a > 0).sum() > 20
    assert np.allclose(img[..., 3][safe], exp_a[safe], atol=2e-4)
    assert np.allclose(img[..., 0][safe], exp_a[safe] * 1.0, atol=2e-4)
    assert np.allclose(img[..., 1][safe], exp_a[safe] * (128 / 255), atol=2e-4)
    assert np.all(img[..., 2] == 0)
    # Q8: isotropic on-axis splat -> normalize(vec2(0,0)) -> NaN basis -> nothing is drawn
    cc[0, 0] = q(32767)
    img, st = orc.render(cs, cc, np.array([0], np.uint32), P, MV, W, H, focal, nthreads=1)
    assert st["n_visible"] == 0 and np.all(img == 0)


def test_render_thread_count_invariant_and_bg(gs, orc):
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 3000, 24, 256, 144)
    o = orc.sort(m, fr.view)
    a, sa = orc.render(cs, cc, o, fr.proj, fr.modelview, 256, 144, fr.focal, nthreads=1)
    b, sb = orc.render(cs, cc, o, fr.proj, fr.modelview, 256, 144, fr.focal, nthreads=5)
    assert np.array_equal(a, b) and sa == sb
    c, _ = orc.render(cs, cc, o, fr.proj, fr.modelview, 256, 144, fr.focal, bg=(1, 0, 0, 1))
    T = 1.0 - a[..., 3]
    assert np.allclose(c[..., 0], a[..., 0] + T, atol=1e-5) and np.allclose(c[..., 3], 1.0, atol=1e-5)
    # informational UNORM8-per-blend emulation stays within a few LSB of the float target on this scene
    q, _ = orc.render(cs, cc, o, fr.proj, fr.modelview, 256, 144, fr.focal, unorm8=True)
    assert np.abs(q - a).max() < 0.1


def test_render_order_matters(gs, orc):
    """Blend is order dependent: reversing the draw order changes the frame (so parity needs the exact order)."""
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 3000, 25, 256, 144)
    o = orc.sort(m, fr.view)
    a, _ = orc.render(cs, cc, o, fr.proj, fr.modelview, 256, 144, fr.focal)
    b, _ = orc.render(cs, cc, o[::-1].copy(), fr.proj, fr.modelview, 256, 144, fr.focal)
    assert np.abs(a - b).max() > 0.05


def test_camera_helpers(gs, orc):
    sc = gs.scenes
    cam = sc.orbit_camera(1920, 1080, 17)
    obj = sc.demo_object()
    proj, mv = orc.camera_matrices(cam.matrixWorld.elements, cam.projectionMatrix.elements, obj.matrixWorld.elements)
    p2 = gs.three_math.get_projection_matrix(cam).elements
    m2 = gs.three_math.get_model_view_matrix(cam, obj).elements
    assert np.array_equal(proj, np.array(p2)) and np.array_equal(mv, np.array(m2))  # same fp64 op order
    # independent check: Y * inv(cam) * obj * Y with numpy
    Y = np.diag([1.0, -1.0, 1.0, 1.0])
    cw = np.array(cam.matrixWorld.elements).reshape(4, 4).T
    ow = np.array(obj.matrixWorld.elements).reshape(4, 4).T
    ref = Y @ np.linalg.inv(cw) @ ow @ Y
    assert np.allclose(mv.reshape(4, 4).T, ref, atol=1e-12)
    cut = sc.demo_cutout()
    w2c = orc.world_to_cutout(cut.matrixWorld.elements, obj.matrixWorld.elements)
    assert np.allclose(w2c.reshape(4, 4).T, np.linalg.inv(np.array(cut.matrixWorld.elements).reshape(4, 4).T) @ ow, atol=1e-12)
    assert np.array_equal(w2c, np.array(gs.three_math.world_to_cutout(cut, obj).elements))


def test_golden_vectors(gs, orc):
    """Committed fixtures (tests/golden/make_golden.py): the oracle must keep reproducing them bit for bit."""
    g = np.load(os.path.join(GOLD, "scene64.npz"))
    cs, cc, m = orc.pack(g["rows"])
    assert np.array_equal(cs, g["center_scale"]) and np.array_equal(cc, g["cov_color"]) and np.array_equal(m[:, 15], g["size_alpha"])
    o = orc.sort(m, g["view"])
    assert np.array_equal(o, g["order"])
    oc = orc.sort(m, g["view"], g["cutout"])
    assert np.array_equal(oc, g["order_cutout"])
    img, _ = orc.render(cs, cc, o, g["proj"], g["modelview"], int(g["width"]), int(g["height"]), float(g["focal"]))
    assert np.abs(img - g["frame"]).max() <= 1e-6  # expf may differ by an ulp between libm builds
    g2 = np.load(os.path.join(GOLD, "scene20k.npz"))
    rows = gs.synth_splats(int(g2["n"]), int(g2["seed"]))
    cs, cc, m = orc.pack(rows)
    assert int(np.bitwise_xor.reduce(cc.reshape(-1))) == int(g2["cov_xor"])
    o = orc.sort(m, g2["view"])
    assert np.array_equal(o, g2["order"])
    img, _ = orc.render(cs, cc, o, g2["proj"], g2["modelview"], int(g2["width"]), int(g2["height"]), float(g2["focal"]))
    assert np.abs(img - g2["frame"].astype(np.float32)).max() <= 2e-3  # frame stored as float16


def test_gl_coverage_check_agrees_with_affine_form(gs, orc):
    """Independent check of the coverage definition: GL interpolates vPosition barycentrically over the two triangles
    of the quad (index.js:52-62,158); the rasters use the affine form d.a2 / d.a1.  Evaluated without a1/a2 (vertex
    positions as the shader emits them in fp32, barycentrics in fp64), the keep/discard decision may differ only on a
    vanishing share of boundary pairs and the Gaussian weight of common pairs only at fp32 rounding level."""
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 20000, 0x5EED0001, 256, 144)
    order = orc.sort(m, fr.view)
    cov = orc.coverage_check(cs, cc, order, fr.proj, fr.modelview, 256, 144, fr.focal, nthreads=4)
    _, st = orc.render(cs, cc, order, fr.proj, fr.modelview, 256, 144, fr.focal, nthreads=4)
    assert cov["pairs_affine"] == st["fragments"] > 100000
    assert cov["pairs_in_quad"] > cov["pairs_affine"]
    assert cov["pairs_differ"] <= max(3, 2e-5 * cov["pairs_affine"])
    assert cov["max_dalpha_common"] <= 1e-4
    assert cov["max_alpha_flipped"] <= np.exp(-4.0) + 1e-6  # a flipped pair sits on the r^2 = 4 boundary


def test_depth_interop_oracle(gs, orc):
    """depthTest:true / depthWrite:false (index.js:179-180): fragments behind foreign geometry are rejected, LEQUAL."""
    rows, cs, cc, m, fr = scene_inputs(gs, orc, 5000, 77, 128, 72)
    order = orc.sort(m, fr.view)
    args = (cs, cc, order, fr.proj, fr.modelview, 128, 72, fr.focal)
    base, st0 = orc.render(*args, bg=(0.2, 0.3, 0.4, 1.0), nthreads=2)
    far, _ = orc.render(*args, bg=(0.2, 0.3, 0.4, 1.0), nthreads=2, depth_in=np.ones((72, 128), np.float32))
    assert np.array_equal(base, far)  # every splat in front of the far plane passes (z/w <= 1 after the clip)
    near, st1 = orc.render(*args, bg=(0.2, 0.3, 0.4, 1.0), nthreads=2, depth_in=np.zeros((72, 128), np.float32))
    assert st1["fragments"] == 0 and np.allclose(near, [0.2, 0.3, 0.4, 1.0])
    p = orc.project(cs, cc, order, fr.proj, fr.modelview, 128, 72, fr.focal)
    zw = (p["zndc"][p["visible"] == 1] * np.float32(0.5) + np.float32(0.5)).astype(np.float32)
    cut = np.float32(np.median(zw))
    half = np.full((72, 128), cut, np.float32)
    half[:, 64:] = 1.0
    mid, st2 = orc.render(*args, bg=(0.2, 0.3, 0.4, 1.0), nthreads=2, depth_in=half)
    assert 0 < st2["fragments"] < st0["fragments"]
    assert np.array_equal(mid[:, 64:], base[:, 64:]) and not np.array_equal(mid[:, :64], base[:, :64])
    # LEQUAL: a splat exactly at the stored depth passes, one ulp behind fails
    for j in np.nonzero(p["visible"] == 1)[0]:
        one = order[j:j + 1]
        _, s1 = orc.render(cs, cc, one, fr.proj, fr.modelview, 128, 72, fr.focal, nthreads=1)
        if s1["fragments"] > 0:
            break
    z = np.float32(p["zndc"][j] * np.float32(0.5) + np.float32(0.5))
    a, sa = orc.render(cs, cc, one, fr.proj, fr.modelview, 128, 72, fr.focal, nthreads=1, depth_in=np.full((72, 128), z, np.float32))
    b, sb = orc.render(cs, cc, one, fr.proj, fr.modelview, 128, 72, fr.focal, nthreads=1,
                       depth_in=np.full((72, 128), np.nextafter(z, np.float32(0)), np.float32))
    assert sa["fragments"] == s1["fragments"] > 0 and sb["fragments"] == 0
