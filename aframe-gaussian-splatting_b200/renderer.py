"""Thin Python object over the C ABI (include/gsplat_b200.h).  One SplatContext == one gs_context ==
one GPU.  No computation happens here: numpy arrays are only the host buffers the ABI reads/writes."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (GS_FORMAT_RGBA8, GS_FORMAT_RGBA32F, GS_RENDER_OUT_DEVICE, GS_RENDER_OUT_TILED,
                   GS_RENDER_REUSE_SORT, GS_RENDER_STATS, GsRenderParams, GsStats)
from .scenes import FrameInputs


class GsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gsplat_b200 error {code}: {msg}")
        self.code = code


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SplatContext:
    """Owner of one gs_context.  Mirrors the worker protocol (clear / push / sort, index.js:572-598) and
    the draw (index.js:184-207)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.gs_create(int(device), C.byref(h))
        if rc != 0:
            raise GsError(rc, (self._lib.gs_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)

    # -- lifetime --
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.gs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise GsError(rc, (self._lib.gs_last_error(self._h) or b"").decode())

    # -- worker protocol --
    def clear(self) -> None:
        self._check(self._lib.gs_clear(self._h))

    def reserve(self, n_total: int) -> None:
        """gs_reserve: size the table for n_total splats (initGL(numVertexes), index.js:248-251)."""
        self._check(self._lib.gs_reserve(self._h, int(n_total)))

    def push_splats(self, rows: np.ndarray) -> None:
        """rows: (n, 32) uint8 raw .splat rows (pushDataBuffer, index.js:328)."""
        rows = np.ascontiguousarray(rows, dtype=np.uint8).reshape(-1, 32)
        self._check(self._lib.gs_push_splats(self._h, _ptr(rows), rows.shape[0]))

    def push_packed(self, center_scale: np.ndarray, cov_color: np.ndarray, size_alpha: np.ndarray) -> None:
        cs = np.ascontiguousarray(center_scale, dtype=np.float32).reshape(-1, 4)
        cc = np.ascontiguousarray(cov_color, dtype=np.uint32).reshape(-1, 4)
        sa = np.ascontiguousarray(size_alpha, dtype=np.float32).reshape(-1)
        if not (cs.shape[0] == cc.shape[0] == sa.shape[0]):
            raise ValueError("packed arrays disagree on the splat count")
        self._check(self._lib.gs_push_packed(self._h, _ptr(cs), _ptr(cc), _ptr(sa), cs.shape[0]))

    @property
    def num_splats(self) -> int:
        n = C.c_uint32()
        self._check(self._lib.gs_num_splats(self._h, C.byref(n)))
        return n.value

    def read_packed(self, first: int = 0, n: Optional[int] = None):
        n = self.num_splats - first if n is None else n
        cs = np.empty((n, 4), np.float32)
        cc = np.empty((n, 4), np.uint32)
        sa = np.empty((n,), np.float32)
        self._check(self._lib.gs_read_packed(self._h, first, n, _ptr(cs), _ptr(cc), _ptr(sa)))
        return cs, cc, sa

    def sort(self, view: np.ndarray, cutout: Optional[np.ndarray] = None, readback: bool = True) -> np.ndarray:
        """{method:'sort'} (index.js:587-596): returns the reference's sortedIndexes (uint32)."""
        v = np.ascontiguousarray(view, dtype=np.float32).reshape(4)
        cu = None if cutout is None else np.ascontiguousarray(cutout, dtype=np.float32).reshape(16)
        out = np.empty((self.num_splats,), np.uint32) if readback else None
        cnt = C.c_uint32()
        self._check(self._lib.gs_sort(self._h, v.ctypes.data_as(C.POINTER(C.c_float)),
                                      None if cu is None else cu.ctypes.data_as(C.POINTER(C.c_float)),
                                      _ptr(out), C.byref(cnt)))
        self.last_sort_count = cnt.value
        return out[:cnt.value] if readback else np.empty((0,), np.uint32)

    # -- draw --
    def make_params(self, frame: FrameInputs, bg=(0.0, 0.0, 0.0, 0.0), fmt: int = GS_FORMAT_RGBA8, flags: int = 0,
                    depth_in: Optional[np.ndarray] = None) -> GsRenderParams:
        """depth_in: optional (H, W) float32 window-space depth of the geometry already drawn (index.js:179-180);
        the returned struct keeps a reference to it."""
        p = GsRenderParams()
        p.proj[:] = [float(x) for x in np.asarray(frame.proj, np.float32).reshape(16)]
        p.modelview[:] = [float(x) for x in np.asarray(frame.modelview, np.float32).reshape(16)]
        p.width, p.height, p.focal = int(frame.width), int(frame.height), float(frame.focal)
        p.bg_rgba[:] = [float(x) for x in bg]
        if frame.cutout is not None:
            p.has_cutout = 1
            p.cutout16[:] = [float(x) for x in np.asarray(frame.cutout, np.float32).reshape(16)]
        p.out_format = fmt
        p.flags = flags
        if depth_in is not None:
            d = np.ascontiguousarray(depth_in, dtype=np.float32)
            if d.size != frame.width * frame.height:
                raise ValueError("depth_in must hold width*height floats")
            p._depth_keepalive = d
            p.depth_in = d.ctypes.data
        return p

    def render(self, frame: FrameInputs, bg=(0.0, 0.0, 0.0, 0.0), fmt: int = GS_FORMAT_RGBA8, out: Optional[np.ndarray] = None,
               reuse_sort: bool = False, depth_in: Optional[np.ndarray] = None, stats: bool = False) -> np.ndarray:
        """One frame into host memory: (H, W, 4) uint8 or float32, row 0 = bottom (GL orientation)."""
        dtype = np.uint8 if fmt == GS_FORMAT_RGBA8 else np.float32
        if out is None:
            out = np.empty((frame.height, frame.width, 4), dtype)
        assert out.dtype == dtype and out.size == frame.height * frame.width * 4 and out.flags["C_CONTIGUOUS"]
        p = self.make_params(frame, bg, fmt, (GS_RENDER_REUSE_SORT if reuse_sort else 0) | (GS_RENDER_STATS if stats else 0),
                             depth_in=depth_in)
        st = GsStats()
        self._check(self._lib.gs_render(self._h, C.byref(p), _ptr(out), C.byref(st)))
        self.last_stats = st
        return out

    def render_stereo(self, view: np.ndarray, eyes, cutout: Optional[np.ndarray] = None, bg=(0.0, 0.0, 0.0, 0.0),
                      fmt: int = GS_FORMAT_RGBA8):
        """gs_render_stereo: one sort with the head camera's `view` (+ cutout), one draw per eye (two FrameInputs).
        Returns the two frames, row 0 = bottom."""
        assert len(eyes) == 2
        dtype = np.uint8 if fmt == GS_FORMAT_RGBA8 else np.float32
        outs = [np.empty((e.height, e.width, 4), dtype) for e in eyes]
        arr = (GsRenderParams * 2)()
        keep = [self.make_params(e, bg, fmt, 0) for e in eyes]
        for i in range(2):
            C.memmove(C.addressof(arr[i]), C.addressof(keep[i]), C.sizeof(GsRenderParams))
        ptrs = (C.c_void_p * 2)(outs[0].ctypes.data, outs[1].ctypes.data)
        stats = (GsStats * 2)()
        v = np.ascontiguousarray(view, dtype=np.float32).reshape(4)
        cu = None if cutout is None else np.ascontiguousarray(cutout, dtype=np.float32).reshape(16)
        self._check(self._lib.gs_render_stereo(self._h, v.ctypes.data_as(C.POINTER(C.c_float)),
                                               None if cu is None else cu.ctypes.data_as(C.POINTER(C.c_float)),
                                               arr, ptrs, stats))
        self.last_stereo_stats = [stats[0], stats[1]]
        return outs

    def render_raw(self, params: GsRenderParams, out_ptr: int) -> GsStats:
        """gs_render with a caller-provided pointer (device pointer when GS_RENDER_OUT_DEVICE is set)."""
        st = GsStats()
        self._check(self._lib.gs_render(self._h, C.byref(params), C.c_void_p(out_ptr), C.byref(st)))
        self.last_stats = st
        return st

    def render_async(self, params: GsRenderParams, out_ptr: int) -> int:
        """gs_render_async: enqueue one frame, return its ticket (four frames may be outstanding: three stages + the copy to the host)."""
        t = C.c_uint64()
        self._check(self._lib.gs_render_async(self._h, C.byref(params), C.c_void_p(out_ptr), C.byref(t)))
        return t.value

    def wait(self, ticket: int) -> GsStats:
        """gs_wait: block until the frame of `ticket` (and its counters) are on the host."""
        st = GsStats()
        self._check(self._lib.gs_wait(self._h, ticket, C.byref(st)))
        self.last_stats = st
        return st

    def read_projected(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.num_splats - first if n is None else n
        out = np.empty((n, 8), np.float32)
        self._check(self._lib.gs_read_projected(self._h, first, n, _ptr(out)))
        return out

    def stats(self) -> dict:
        st = GsStats()
        self._check(self._lib.gs_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    # -- multi-GPU --
    def set_shard(self, rank: int, world: int) -> None:
        self._check(self._lib.gs_set_shard(self._h, rank, world))

    def owned_tiles(self, width: int, height: int, rank: int, world: int) -> int:
        return int(self._lib.gs_owned_tiles(width, height, rank, world))

    def assemble_tiles(self, gathered_ptr: int, tiles_per_rank: int, world: int, width: int, height: int, fmt: int, out_ptr: int) -> None:
        self._check(self._lib.gs_assemble_tiles(self._h, C.c_void_p(gathered_ptr), tiles_per_rank, world, width, height, fmt,
                                                C.c_void_p(out_ptr)))

    def peer_export(self, frame_bytes: int) -> bytes:
        """gs_peer_export: allocate this rank's shared frame ring, return its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.gs_peer_export(self._h, frame_bytes, buf))
        return buf.raw

    def peer_import(self, rank: int, world: int, handles: list) -> None:
        """gs_peer_import: map every rank's ring (handles in rank order, each 64 bytes)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._check(self._lib.gs_peer_import(self._h, rank, world, C.c_char_p(blob)))

    def peer_frame(self, ticket: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.gs_peer_frame(self._h, ticket, C.byref(p)))
        return p.value

    # -- memory helpers --
    def host_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.gs_host_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def host_free(self, ptr: int) -> None:
        self._check(self._lib.gs_host_free(self._h, C.c_void_p(ptr)))

    def pinned_array(self, shape, dtype) -> np.ndarray:
        """numpy view over page-locked memory owned by the context (freed with the context's process)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self.host_alloc(nbytes)
        buf = (C.c_uint8 * nbytes).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.gs_device_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr: int) -> None:
        self._check(self._lib.gs_device_free(self._h, C.c_void_p(ptr)))

    def memcpy_d2h(self, dst: np.ndarray, src_ptr: int, nbytes: int) -> None:
        self._check(self._lib.gs_memcpy_d2h(self._h, _ptr(dst), C.c_void_p(src_ptr), nbytes))

    def synchronize(self) -> None:
        self._check(self._lib.gs_synchronize(self._h))
