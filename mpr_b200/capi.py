"""ctypes bindings for libmprb.so (the C ABI declared in include/mprb.h).

This is harness glue for tests and bench.py: the product is the shared
library.  Nothing here computes; every call goes into the CUDA path, and the
loader raises if the library has not been built (there is no fallback).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
# MPRB_LIBRARY: another build of the library (A/B timing of two builds in one job); default = the in-tree one
LIB_PATH = Path(os.environ["MPRB_LIBRARY"]) if os.environ.get("MPRB_LIBRARY") else _PKG / "libmprb.so"


class TileNode(C.Structure):
    _fields_ = [("position", C.c_int32), ("tape", C.c_int32), ("next", C.c_int32)]


TILE_DTYPE = np.dtype([("position", "<i4"), ("tape", "<i4"), ("next", "<i4")])


class Buffers(C.Structure):
    _fields_ = [
        ("image_size_px", C.c_int32),
        ("filled", C.POINTER(C.c_int32) * 4),
        ("tiles", C.POINTER(TileNode) * 4),
        ("tile_array_size", C.c_uint64 * 4),
        ("tape_data", C.POINTER(C.c_uint64)),
        ("tape_index", C.POINTER(C.c_int32)),
        ("num_active_tiles", C.POINTER(C.c_int32)),
        ("normals", C.POINTER(C.c_uint32)),
    ]


class CtxOpts(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("num_subtapes", C.c_int64),
        ("row_begin", C.c_int32),
        ("row_end", C.c_int32),
        ("row_mod", C.c_int32),
        ("row_rem", C.c_int32),
        ("col_step", C.c_int32),
        ("n_gpus", C.c_int32),
    ]


class FrameStats(C.Structure):
    _fields_ = [
        ("n_active", C.c_int32 * 3),
        ("tape_index", C.c_int32),
        ("overflow", C.c_int32),
        ("i_tiles", C.c_uint64 * 3),
        ("i_cells", C.c_uint64 * 3),
        ("p_tiles", C.c_uint64 * 3),
        ("p_cells", C.c_uint64 * 3),
        ("p_kept", C.c_uint64 * 3),
        ("f_tiles", C.c_uint64),
        ("f_cells", C.c_uint64),
        ("n_pixels", C.c_uint64),
        ("n_cells", C.c_uint64),
        ("gpu_ms", C.c_float),
        ("kernel_ms", C.c_float * 12),
        ("n_launches", C.c_int32),
        ("f_items", C.c_uint64),
        ("p_written", C.c_uint64),
        ("i_sub_tiles", C.c_uint64),
    ]

    def asdict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


EXPORTS = [
    "mprb_ctx_create", "mprb_ctx_destroy", "mprb_ctx_buffers", "mprb_ctx_set_timing",
    "mprb_tape_create", "mprb_tape_destroy", "mprb_tape_data", "mprb_tape_length",
    "mprb_tape_num_slots", "mprb_render2d", "mprb_render3d", "mprb_render2d_host",
    "mprb_render3d_host", "mprb_frame_stats_get", "mprb_tape_from_frep", "mprb_free", "mprb_free_device",
    "mprb_last_error", "mprb_version", "mprb_effects_create", "mprb_effects_destroy",
    "mprb_effects_draw_ssao", "mprb_effects_draw_shaded", "mprb_effects_buffers",
    "mprb_malloc_managed", "mprb_exchange_bytes", "mprb_exchange_pack", "mprb_exchange_unpack", "mprb_ctx_publish", "mprb_render2d_brute", "mprb_render2d_heatmap", "mprb_render3d_heatmap",
]

_lib = None


def lib():
    """Loads libmprb.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
            "there is no CPU fallback for the render path")
    L = C.CDLL(str(LIB_PATH))   # RTLD_LOCAL: the C++ mpr:: symbols inside must not leak
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    L.mprb_ctx_create.argtypes = [i32, C.POINTER(CtxOpts), C.POINTER(vp)]
    L.mprb_ctx_destroy.argtypes = [vp]
    L.mprb_ctx_destroy.restype = None
    L.mprb_ctx_buffers.argtypes = [vp, C.POINTER(Buffers)]
    L.mprb_ctx_set_timing.argtypes = [vp, C.c_int]
    L.mprb_tape_create.argtypes = [C.c_void_p, i32, C.POINTER(vp)]
    L.mprb_tape_destroy.argtypes = [vp]
    L.mprb_tape_destroy.restype = None
    L.mprb_tape_data.argtypes = [vp]
    L.mprb_tape_data.restype = C.c_void_p
    L.mprb_tape_length.argtypes = [vp]
    L.mprb_tape_length.restype = i32
    L.mprb_tape_num_slots.argtypes = [vp]
    L.mprb_tape_num_slots.restype = i32
    L.mprb_render2d.argtypes = [vp, vp, C.POINTER(f32), f32]
    L.mprb_render3d.argtypes = [vp, vp, C.POINTER(f32)]
    L.mprb_render2d_host.argtypes = [vp, C.c_void_p, i32, C.POINTER(f32), f32, C.c_void_p]
    L.mprb_render3d_host.argtypes = [vp, C.c_void_p, i32, C.POINTER(f32), C.c_void_p, C.c_void_p]
    L.mprb_frame_stats_get.argtypes = [vp, C.POINTER(FrameStats)]
    L.mprb_tape_from_frep.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.POINTER(C.c_uint64)),
                                      C.POINTER(i32), C.POINTER(i32)]
    L.mprb_free.argtypes = [C.c_void_p]
    L.mprb_free.restype = None
    L.mprb_render2d_brute.argtypes = [vp, vp, C.c_void_p, C.c_float]
    L.mprb_render2d_heatmap.argtypes = [vp, vp, C.c_void_p, C.c_float, C.POINTER(C.POINTER(C.c_float))]
    L.mprb_render3d_heatmap.argtypes = [vp, vp, C.c_void_p, C.POINTER(C.POINTER(C.c_float))]
    L.mprb_exchange_bytes.argtypes = [vp, C.c_int]
    L.mprb_exchange_bytes.restype = C.c_size_t
    L.mprb_exchange_pack.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p]
    L.mprb_exchange_unpack.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p]
    if hasattr(L, "mprb_ctx_publish"):          # absent from older builds loaded through MPRB_LIBRARY
        L.mprb_ctx_publish.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p]
    L.mprb_effects_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(vp)]
    L.mprb_effects_destroy.argtypes = [vp]
    L.mprb_effects_destroy.restype = None
    L.mprb_effects_draw_ssao.argtypes = [vp, vp]
    L.mprb_effects_draw_shaded.argtypes = [vp, vp]
    L.mprb_effects_buffers.argtypes = [vp, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32))]
    L.mprb_last_error.restype = C.c_char_p
    L.mprb_version.restype = C.c_char_p
    _lib = L
    return L


class MprbError(RuntimeError):
    pass


def _check(code):
    if code != 0:
        raise MprbError(f"mprb error {code}: {lib().mprb_last_error().decode()}")


def tape_from_frep(data: bytes, simplify: bool = True) -> np.ndarray:
    """.frep bytes -> packed uint64 tape (host only; no GPU needed)."""
    cells = C.POINTER(C.c_uint64)()
    n = C.c_int32()
    ns = C.c_int32()
    _check(lib().mprb_tape_from_frep(data, len(data), int(simplify), C.byref(cells), C.byref(n), C.byref(ns)))
    out = np.ctypeslib.as_array(cells, shape=(n.value,)).copy()
    lib().mprb_free(cells)
    return out


def mat_colmajor(m) -> np.ndarray:
    """Row-major nested matrix (as written in maths) -> column-major float32 vector."""
    a = np.asarray(m, dtype=np.float32)
    return np.ascontiguousarray(a.T).reshape(-1)


def view_matrix_3d() -> np.ndarray:
    """T = I with T(3,2) = 0.3, the perspective the reference's 3D drivers use
    (reference benchmark/render_3d_table.cpp:48-49)."""
    t = np.eye(4, dtype=np.float32)
    t[3, 2] = 0.3
    return t


class Tape:
    """Mirror of mpr::Tape (reference inc/tape.hpp:24-30): packed cells in GPU-visible memory."""

    def __init__(self, cells: np.ndarray):
        self.cells = np.ascontiguousarray(cells, dtype=np.uint64)
        self._h = C.c_void_p()
        _check(lib().mprb_tape_create(self.cells.ctypes.data, self.cells.size, C.byref(self._h)))
        self.length = lib().mprb_tape_length(self._h)
        self.num_slots = lib().mprb_tape_num_slots(self._h)

    def close(self):
        if self._h:
            lib().mprb_tape_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def glibc_effect_samples():
    """The SSAO sample sets exactly as mpr::Effects::Effects() draws them (src/effects.cu:229-250):
    glibc rand() from its default seed, 64x3 kernel then 256x3 rotation vectors, column-major."""
    libc = C.CDLL("libc.so.6")
    libc.srand(1)
    rmax = 2147483647.0
    def r():
        return np.float32(np.float32(libc.rand()) / np.float32(rmax))
    kernel = np.zeros((64, 3), dtype=np.float32)
    for i in range(64):
        v = np.array([np.float32(2.0) * (r() - np.float32(0.5)), np.float32(2.0) * (r() - np.float32(0.5)), r()],
                     dtype=np.float32)
        v = v / np.float32(np.sqrt(np.float32((v * v).sum(dtype=np.float32))))
        scale = np.float32(i) / np.float32(63)
        scale = (scale * scale) * np.float32(0.9) + np.float32(0.1)
        kernel[i] = v * scale
    rvecs = np.zeros((256, 3), dtype=np.float32)
    for i in range(256):
        v = np.array([np.float32(2.0) * (r() - np.float32(0.5)), np.float32(2.0) * (r() - np.float32(0.5)), 0.0],
                     dtype=np.float32)
        rvecs[i] = v / np.float32(np.sqrt(np.float32((v * v).sum(dtype=np.float32))))
    return np.asfortranarray(kernel), np.asfortranarray(rvecs)


class Effects:
    """Mirror of mpr::Effects (reference inc/effects.hpp:21-37)."""

    def __init__(self, kernel=None, rvecs=None):
        if kernel is None:
            kernel, rvecs = glibc_effect_samples()
        self.kernel = np.asfortranarray(kernel, dtype=np.float32)
        self.rvecs = np.asfortranarray(rvecs, dtype=np.float32)
        self._h = C.c_void_p()
        _check(lib().mprb_effects_create(self.kernel.ctypes.data, self.rvecs.ctypes.data, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mprb_effects_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def drawSSAO(self, ctx):
        _check(lib().mprb_effects_draw_ssao(self._h, ctx._h))
        self._size = ctx.image_size_px

    def drawShaded(self, ctx):
        _check(lib().mprb_effects_draw_shaded(self._h, ctx._h))
        self._size = ctx.image_size_px

    def image(self):
        img = C.POINTER(C.c_int32)()
        _check(lib().mprb_effects_buffers(self._h, C.byref(img), None))
        return np.ctypeslib.as_array(img, shape=(self._size, self._size))

    def tmp(self):
        tmp = C.POINTER(C.c_int32)()
        _check(lib().mprb_effects_buffers(self._h, None, C.byref(tmp)))
        return np.ctypeslib.as_array(tmp, shape=(self._size, self._size))


class Context:
    """Mirror of mpr::Context (reference inc/context.hpp:38-73)."""

    def __init__(self, image_size_px: int, device: int = -1, num_subtapes: int = 0,
                 row_begin: int = 0, row_end: int = 0, row_mod: int = 0, row_rem: int = 0, col_step: int = 0,
                 n_gpus: int = 1):
        """n_gpus > 1: one context over that many devices of this process (see include/mprb.h);
        n_gpus = 0 takes the count from MPRB_GPUS, like mpr::Context does."""
        self.image_size_px = image_size_px
        self._h = C.c_void_p()
        opts = CtxOpts(device, num_subtapes, row_begin, row_end, row_mod, row_rem, col_step, n_gpus)
        _check(lib().mprb_ctx_create(image_size_px, C.byref(opts), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mprb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_timing(self, on: bool):
        _check(lib().mprb_ctx_set_timing(self._h, int(on)))

    def render2D(self, tape: Tape, mat=None, z: float = 0.0):
        m = mat_colmajor(np.eye(3) if mat is None else mat)
        _check(lib().mprb_render2d(self._h, tape._h, m.ctypes.data_as(C.POINTER(C.c_float)), z))

    def render3D(self, tape: Tape, mat=None):
        m = mat_colmajor(view_matrix_3d() if mat is None else mat)
        _check(lib().mprb_render3d(self._h, tape._h, m.ctypes.data_as(C.POINTER(C.c_float))))

    def render2D_brute(self, tape: Tape, mat=None, z: float = 0.0):
        m = mat_colmajor(np.eye(3) if mat is None else mat)
        _check(lib().mprb_render2d_brute(self._h, tape._h, m.ctypes.data, z))

    def _take_heatmap(self, ptr):
        s = self.image_size_px
        out = np.ctypeslib.as_array(ptr, shape=(s, s)).copy()
        lib().mprb_free_device(ptr)
        return out

    def render2D_heatmap(self, tape: Tape, mat=None, z: float = 0.0):
        m = mat_colmajor(np.eye(3) if mat is None else mat)
        h = C.POINTER(C.c_float)()
        _check(lib().mprb_render2d_heatmap(self._h, tape._h, m.ctypes.data, z, C.byref(h)))
        return self._take_heatmap(h)

    def render3D_heatmap(self, tape: Tape, mat=None):
        m = mat_colmajor(view_matrix_3d() if mat is None else mat)
        h = C.POINTER(C.c_float)()
        _check(lib().mprb_render3d_heatmap(self._h, tape._h, m.ctypes.data, C.byref(h)))
        return self._take_heatmap(h)

    def exchange_bytes(self, dim: int) -> int:
        return int(lib().mprb_exchange_bytes(self._h, dim))

    def exchange_pack(self, dim: int, dst_ptr: int, stream: int = 0):
        _check(lib().mprb_exchange_pack(self._h, dim, dst_ptr, stream))

    def exchange_unpack(self, dim: int, src_ptr: int, stream: int = 0):
        _check(lib().mprb_exchange_unpack(self._h, dim, src_ptr, stream))

    def publish(self, dim: int, image_ptr: int, normals_ptr: int = 0):
        """Owned 64x64-px blocks of the last frame -> full-size images at these addresses (page-locked
        host memory or a peer device's); mprb_ctx_publish."""
        _check(lib().mprb_ctx_publish(self._h, dim, image_ptr, normals_ptr or None))

    def render2D_host(self, cells: np.ndarray, image_out: np.ndarray, mat=None, z: float = 0.0):
        m = mat_colmajor(np.eye(3) if mat is None else mat)
        _check(lib().mprb_render2d_host(self._h, cells.ctypes.data, cells.size,
                                        m.ctypes.data_as(C.POINTER(C.c_float)), z,
                                        image_out.ctypes.data if image_out is not None else None))

    def render3D_host(self, cells: np.ndarray, depth_out: np.ndarray, normals_out=None, mat=None):
        m = mat_colmajor(view_matrix_3d() if mat is None else mat)
        _check(lib().mprb_render3d_host(self._h, cells.ctypes.data, cells.size,
                                        m.ctypes.data_as(C.POINTER(C.c_float)),
                                        depth_out.ctypes.data if depth_out is not None else None,
                                        normals_out.ctypes.data if normals_out is not None else None))

    # -- buffer views (managed memory; valid until the next render call) ----------
    def buffers(self) -> Buffers:
        b = Buffers()
        _check(lib().mprb_ctx_buffers(self._h, C.byref(b)))
        return b

    def stats(self) -> FrameStats:
        s = FrameStats()
        _check(lib().mprb_frame_stats_get(self._h, C.byref(s)))
        return s

    def image(self) -> np.ndarray:
        """stages[3].filled as an (S, S) int32 array indexed [y, x]."""
        b = self.buffers()
        s = self.image_size_px
        return np.ctypeslib.as_array(b.filled[3], shape=(s, s))

    def filled(self, stage: int) -> np.ndarray:
        b = self.buffers()
        side = self.image_size_px // (64 >> (2 * stage))
        return np.ctypeslib.as_array(b.filled[stage], shape=(side, side))

    def normals(self) -> np.ndarray:
        b = self.buffers()
        s = self.image_size_px
        return np.ctypeslib.as_array(b.normals, shape=(s, s))

    def tiles(self, stage: int) -> np.ndarray:
        b = self.buffers()
        n = int(b.tile_array_size[stage])
        if n == 0:
            return np.zeros(0, dtype=TILE_DTYPE)
        raw = np.ctypeslib.as_array(C.cast(b.tiles[stage], C.POINTER(C.c_int32)), shape=(n, 3))
        return raw.view(TILE_DTYPE).reshape(n)

    def device_image(self):
        """(ptr, nbytes) of stages[3].filled for wrapping as a device array."""
        b = self.buffers()
        return C.cast(b.filled[3], C.c_void_p).value, self.image_size_px * self.image_size_px * 4

    def device_normals(self):
        b = self.buffers()
        return C.cast(b.normals, C.c_void_p).value, self.image_size_px * self.image_size_px * 4

    def tape_data(self, n_cells=None) -> np.ndarray:
        b = self.buffers()
        n = int(b.tape_index[0]) if n_cells is None else n_cells
        return np.ctypeslib.as_array(b.tape_data, shape=(n,))
