"""aframe-gaussian-splatting_b200 — B200-native sort + splat-raster path behind the reference component's API.

The directory name carries a hyphen, so import it with
    import importlib; gs = importlib.import_module("aframe-gaussian-splatting_b200")
or through the root-level alias module `aframe_gaussian_splatting_b200`.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI), the ctypes binding, the
host-side mirror of the reference's component interface, and the synthetic scene generator.
"""
from . import _lib, build, dist, ply, scenes, three_math  # noqa: F401
from ._lib import (GS_FORMAT_RGBA8, GS_FORMAT_RGBA32F, GS_RENDER_OUT_DEVICE, GS_RENDER_OUT_PEER,  # noqa: F401
                   GS_RENDER_OUT_TILED, GS_RENDER_REUSE_SORT, GS_RENDER_STATS, GS_RENDER_DEPTH_DEVICE, GsRenderParams,
                   GsStats)
from .renderer import GsError, SplatContext  # noqa: F401
from .scenes import FrameInputs, make_frame, synth_splats  # noqa: F401
from .component import GaussianSplattingComponent, SortWorker  # noqa: F401

__all__ = ["SplatContext", "GsError", "GaussianSplattingComponent", "SortWorker", "FrameInputs", "make_frame",
           "synth_splats", "scenes", "three_math", "build"]
