"""Host-side mirror of the reference's plugin interface for the hot path.

The reference registers one A-Frame component, `gaussian_splatting` (index.js:1), whose public surface is
the schema (`src`, `cutoutEntity`, `pixelRatio`, `xrPixelRatio`, index.js:2-7), the vanilla-three entry
`loadData(camera, object, renderer, src)` (index.js:24,222) and the per-frame `tick()` (index.js:438).
Node / A-Frame are not available in this image, so the same interface is mirrored here in Python with the
same names, argument meaning and error behaviour; every method forwards to the C ABI
(include/gsplat_b200.h).  INTEGRATION.md shows the N-API stub that binds the same ABI from JavaScript.

What disappears on the GPU: the worker's `sortedIndexes` never leave the device unless asked for
(`worker.onmessage`, index.js:201-207, becomes a resident order), and the texture upload of
`pushDataBuffer` (index.js:404-431) is the device-side pack itself.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Optional

import numpy as np

from . import ply as _ply
from .renderer import SplatContext
from .scenes import FrameInputs
from .three_math import (Matrix4, Object3D, PerspectiveCamera, focal_length, get_model_view_matrix,
                         get_projection_matrix, world_to_cutout)
from ._lib import GS_FORMAT_RGBA8

ROW_LENGTH = 3 * 4 + 3 * 4 + 4 + 4  # index.js:227


class SortWorker:
    """The Web Worker's message protocol (index.js:572-598) served by the GPU context.

    postMessage({'method': 'clear'})                                    -> None
    postMessage({'method': 'push', 'rows': uint8[n*32]})                -> None   (raw rows; the pack that the
        reference runs on the main thread before posting `matrices` happens on the device)
    postMessage({'method': 'sort', 'view': f32[4], 'cutout': f32[16]?}) -> {'sortedIndexes': uint32[V]}
    `onmessage`, when set, receives the reply like the main thread's handler (index.js:201).
    """

    def __init__(self, ctx: SplatContext):
        self.ctx = ctx
        self.onmessage: Optional[Callable[[dict], None]] = None

    def postMessage(self, data: dict, readback: bool = True):
        method = data.get("method")
        if method == "clear":
            self.ctx.clear()
            return None
        if method == "push":
            self.ctx.push_splats(np.frombuffer(memoryview(data["rows"]), dtype=np.uint8))
            return None
        if method == "sort":
            if self.ctx.num_splats == 0:
                # index.js:588-590 replies Uint32Array(1) == [0] (quirk Q7: one garbage instance)
                reply = {"sortedIndexes": np.zeros(1, np.uint32)}
            else:
                reply = {"sortedIndexes": self.ctx.sort(data["view"], data.get("cutout"), readback=readback)}
            if self.onmessage is not None:
                self.onmessage(reply)
            return reply
        return None  # unknown methods are ignored, as in the reference


class GaussianSplattingComponent:
    """`gaussian_splatting` (index.js:1-746) for the sort + draw path."""

    schema = {  # index.js:2-7
        "src": {"type": "string", "default": "train.splat"},
        "cutoutEntity": {"type": "selector"},
        "pixelRatio": {"type": "number", "default": 1},
        "xrPixelRatio": {"type": "number", "default": 0.5},
    }

    def __init__(self, data: Optional[dict] = None, device: int = 0):
        self.data = {k: v.get("default") for k, v in self.schema.items()}
        self.data.update(data or {})
        self.device = device
        self.cutout: Optional[Object3D] = None
        self.camera: Optional[PerspectiveCamera] = None
        self.object: Optional[Object3D] = None
        self.renderer: Optional[SplatContext] = None
        self.worker: Optional[SortWorker] = None
        self.loadedVertexCount = 0
        self.rowLength = ROW_LENGTH
        self.sortReady = False
        self.instanceCount = 0
        self.pixelRatio = 1.0
        self._have_order = False

    # ---- index.js:8-23 ----
    def init(self, camera: PerspectiveCamera, object3d: Object3D, renderer: Optional[SplatContext] = None):
        if self.data["pixelRatio"] and self.data["pixelRatio"] > 0:
            self.pixelRatio = float(self.data["pixelRatio"])  # renderer.setPixelRatio
        renderer = renderer or SplatContext(self.device)
        self.loadData(camera, object3d, renderer, self.data["src"])
        if self.data.get("cutoutEntity") is not None:
            self.cutout = self.data["cutoutEntity"]

    # ---- index.js:25-221 ----
    def initGL(self, numVertexes: int) -> None:
        """The reference sizes its two data textures from numVertexes (index.js:26-46, known from the Content-Length,
        index.js:248-251); here the resident table is reserved for as many splats, so the pushes that follow never
        have to grow it (and never wait for frames in flight).  sortReady flips exactly as at index.js:220."""
        if numVertexes > 0 and self.renderer is not None:
            self.renderer.reserve(int(numVertexes))
        self.sortReady = True

    # ---- index.js:222-327 ----
    def loadData(self, camera, object3d, renderer: SplatContext, src) -> None:
        self.camera, self.object, self.renderer = camera, object3d, renderer
        self.loadedVertexCount = 0
        self.worker = SortWorker(renderer)
        self.worker.onmessage = self._on_sorted
        self.worker.postMessage({"method": "clear"})
        if isinstance(src, (bytes, bytearray, memoryview, np.ndarray)):
            buf, is_ply = np.frombuffer(memoryview(src), dtype=np.uint8), False
        else:
            is_ply = str(src).endswith(".ply")  # index.js:257
            with open(os.fspath(src), "rb") as f:
                buf = np.frombuffer(f.read(), dtype=np.uint8)
        if is_ply:
            # a .ply is converted first and sized afterwards, the reference's path when no Content-Length is known
            # (index.js:315-323); sizing from the raw byte count (index.js:249-250) would over-reserve 248/32 x (Q11)
            buf = np.frombuffer(self.processPlyBuffer(buf.tobytes()), dtype=np.uint8)  # index.js:315-317
        self.initGL(len(buf) // self.rowLength)  # index.js:249-250 / 320-323
        # progressive push in chunks, whole rows only (index.js:279-298); a trailing partial row is dropped
        n_rows = len(buf) // self.rowLength
        chunk = 1 << 22
        for first in range(0, n_rows, chunk):
            cnt = min(chunk, n_rows - first)
            self.pushDataBuffer(buf[first * self.rowLength:(first + cnt) * self.rowLength], cnt)

    # ---- index.js:328-437 ----
    def pushDataBuffer(self, buffer, vertexCount: int) -> None:
        if vertexCount <= 0:
            return
        rows = np.frombuffer(memoryview(buffer), dtype=np.uint8)[: vertexCount * self.rowLength]
        self.worker.postMessage({"method": "push", "rows": rows})
        self.loadedVertexCount += vertexCount
        self._have_order = False

    # ---- index.js:201-207 ----
    def _on_sorted(self, reply: dict) -> None:
        self.instanceCount = int(getattr(self.renderer, "last_sort_count", len(reply["sortedIndexes"])))
        self.sortReady = True
        self._have_order = True

    # ---- index.js:438-455 ----
    def tick(self, time: float = 0.0, timeDelta: float = 0.0, readback: bool = False):
        if not self.sortReady:
            return None
        self.sortReady = False
        camera_mtx = self.getModelViewMatrix().elements
        view = np.array([camera_mtx[2], camera_mtx[6], camera_mtx[10], camera_mtx[14]], dtype=np.float32)
        cutout = None
        if self.cutout is not None:
            cutout = np.asarray(world_to_cutout(self.cutout, self.object).elements, dtype=np.float32)
        return self.worker.postMessage({"method": "sort", "view": view, "cutout": cutout}, readback=readback)

    # ---- index.js:456-487 ----
    def getProjectionMatrix(self, camera=None) -> Matrix4:
        return get_projection_matrix(camera or self.camera)

    def getModelViewMatrix(self, camera=None) -> Matrix4:
        return get_model_view_matrix(camera or self.camera, self.object)

    # ---- index.js:184-195 + the instanced draw ----
    def _frame_inputs_px(self, w: int, h: int, camera=None) -> FrameInputs:
        """onBeforeRender (index.js:184-195) for a viewport of w x h device pixels."""
        proj = self.getProjectionMatrix(camera)
        mv = self.getModelViewMatrix(camera)
        cut = None
        if self.cutout is not None:
            cut = np.asarray(world_to_cutout(self.cutout, self.object).elements, dtype=np.float32)
        return FrameInputs(proj=np.asarray(proj.elements, np.float32), modelview=np.asarray(mv.elements, np.float32),
                           view=np.array([mv.elements[2], mv.elements[6], mv.elements[10], mv.elements[14]], np.float32),
                           width=w, height=h, focal=float(np.float32(focal_length(h, proj))), cutout=cut)

    def frame_inputs(self, width: int, height: int, camera=None) -> FrameInputs:
        # renderer.setPixelRatio: three.js floors the drawing-buffer size and the current viewport
        # (Math.floor(width * pixelRatio), viewport.multiplyScalar(pixelRatio).floor())
        w, h = int(math.floor(width * self.pixelRatio)), int(math.floor(height * self.pixelRatio))
        return self._frame_inputs_px(w, h, camera)

    def render_xr(self, eye_cameras, width: int, height: int, bg=(0.0, 0.0, 0.0, 0.0), fmt: int = GS_FORMAT_RGBA8):
        """WebXR presentation (index.js:13-15, 184-195).  `init` hands `xrPixelRatio` to
        renderer.xr.setFramebufferScaleFactor, so each eye's viewport is the XR layer's native eye size
        (width x height) scaled by it (floored here).  The frame's single sort request comes from tick(), i.e. from
        `this.camera` - the head pose - (index.js:438-455), while material.onBeforeRender runs once per eye camera
        with that eye's matrices and viewport.  Returns [left, right] frames, row 0 = bottom."""
        ratio = float(self.data.get("xrPixelRatio") or 0)
        if ratio <= 0:
            ratio = 1.0
        w, h = int(math.floor(width * ratio)), int(math.floor(height * ratio))
        eyes = [self._frame_inputs_px(w, h, cam) for cam in eye_cameras]
        head = self.getModelViewMatrix().elements
        view = np.array([head[2], head[6], head[10], head[14]], dtype=np.float32)  # index.js:441-442
        cut = eyes[0].cutout
        frames = self.renderer.render_stereo(view, eyes, cutout=cut, bg=bg, fmt=fmt)
        self._have_order = True
        return frames

    def render(self, width: int, height: int, camera=None, bg=(0.0, 0.0, 0.0, 0.0), fmt: int = GS_FORMAT_RGBA8,
               out: Optional[np.ndarray] = None, synchronous: bool = True) -> np.ndarray:
        """Draw the mesh into an RGBA frame (row 0 = bottom).  synchronous=True sorts with this frame's camera
        (the oracle's definition); synchronous=False draws with the order of the last tick(), which is what the
        reference does while a sort is in flight (index.js:206,439-440)."""
        fr = self.frame_inputs(width, height, camera)
        reuse = (not synchronous) and self._have_order
        return self.renderer.render(fr, bg=bg, fmt=fmt, out=out, reuse_sort=reuse)

    # ---- index.js:600-745 ----
    def processPlyBuffer(self, inputBuffer: bytes) -> bytes:
        return _ply.process_ply_buffer(inputBuffer)
