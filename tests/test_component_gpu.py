"""The host-side mirror of the reference component (component.py) end to end on the GPU: the same call sequence a
page makes (init -> loadData -> pushDataBuffer chunks -> tick -> draw), checked against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(gs):
    sc = gs.scenes
    return sc.fixed_camera(640, 360), sc.demo_object()


def test_component_splat_file_tick_and_draw(gs, orc, tmp_path):
    rows = gs.synth_splats(30000, 77)
    path = tmp_path / "scene.splat"
    path.write_bytes(rows.tobytes() + b"\x01\x02\x03")  # trailing partial row is dropped (index.js:279-298)
    cam, obj = _scene(gs)
    comp = gs.GaussianSplattingComponent({"src": str(path)})
    comp.init(cam, obj)
    assert comp.loadedVertexCount == len(rows) and comp.renderer.num_splats == len(rows) and comp.sortReady
    cs, cc, m = orc.pack(rows)
    fr = gs.make_frame(cam, obj, 640, 360)
    # tick(): the worker's sortedIndexes (readback on request) are the oracle's
    reply = comp.tick(readback=True)
    assert np.array_equal(reply["sortedIndexes"], orc.sort(m, fr.view))
    assert comp.sortReady and comp.instanceCount == len(reply["sortedIndexes"])
    # draw with the order of that tick (reference behaviour) and synchronously: same camera -> same frame
    a = comp.render(640, 360, fmt=gs.GS_FORMAT_RGBA32F, synchronous=False)
    b = comp.render(640, 360, fmt=gs.GS_FORMAT_RGBA32F, synchronous=True)
    exp, _ = orc.render(cs, cc, orc.sort(m, fr.view), fr.proj, fr.modelview, 640, 360, fr.focal)
    assert np.abs(a - exp).max() <= 1e-3 and np.abs(b - exp).max() <= 1e-3
    # matrices are the reference's (index.js:456-487)
    assert np.array_equal(np.asarray(comp.getProjectionMatrix().elements, np.float32), fr.proj)
    assert np.array_equal(np.asarray(comp.getModelViewMatrix().elements, np.float32), fr.modelview)
    comp.renderer.close()


def test_component_ply_cutout_and_pixel_ratio(gs, orc, tmp_path):
    rng = np.random.default_rng(3)
    n = 8000
    xyz = rng.uniform([-2, -1, -3], [2, 2, 1], size=(n, 3)).astype(np.float32)
    blob = gs.ply.write_inria_ply(str(tmp_path / "scene.ply"), xyz, rng.normal(0, 1.2, (n, 3)).astype(np.float32),
                                  rng.normal(1, 2, n).astype(np.float32), rng.normal(-3.5, 0.7, (n, 3)).astype(np.float32),
                                  rng.normal(size=(n, 4)).astype(np.float32))
    cam, obj = _scene(gs)
    cut = gs.three_math.Object3D(position=gs.scenes.DEMO_OBJECT_POSITION, scale=(3.0, 2.0, 3.0))
    comp = gs.GaussianSplattingComponent({"src": str(tmp_path / "scene.ply"), "cutoutEntity": cut, "pixelRatio": 0.5})
    comp.init(cam, obj)
    assert comp.loadedVertexCount == n
    rows = np.frombuffer(comp.processPlyBuffer(blob), np.uint8).reshape(-1, 32)
    cs, cc, m = orc.pack(rows)
    frame = comp.render(640, 360, fmt=gs.GS_FORMAT_RGBA32F)
    assert frame.shape == (180, 320, 4)  # pixelRatio scales the drawing buffer (index.js:10-12)
    fr = comp.frame_inputs(640, 360)
    assert fr.cutout is not None
    order = orc.sort(m, fr.view, fr.cutout)
    assert 0 < len(order) < n
    exp, _ = orc.render(cs, cc, order, fr.proj, fr.modelview, 320, 180, fr.focal)
    assert np.abs(frame - exp).max() <= 1e-3
    # worker protocol: clear drops everything; sort before any push answers [0] like index.js:588-590
    comp.worker.postMessage({"method": "clear"})
    assert comp.renderer.num_splats == 0
    assert np.array_equal(comp.worker.postMessage({"method": "sort", "view": fr.view})["sortedIndexes"], [0])
    comp.renderer.close()
