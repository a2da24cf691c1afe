"""TEST INFRASTRUCTURE ONLY.

ctypes loaders for
  * oracle/_build/libmpr_oracle.so -- the plain-C CPU restatement (mpr_oracle.c), and
  * oracle/_ref/libmpr_ref.so      -- the unmodified reference CUDA renderer behind a C shim
                                      (ref_wrap.cu), usable only on a GPU box.

Only tests/, tools/gpu_check.py, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs import this package; mpr_b200/ never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
ORACLE_SO = _DIR / "_build" / "libmpr_oracle.so"
REF_SO = _DIR / "_ref" / "libmpr_ref.so"

TILE_DTYPE = np.dtype([("position", "<i4"), ("tape", "<i4"), ("next", "<i4")])

_oracle = None
_ref = None


def build(ref: bool = True):
    """(Re)builds the oracle libraries via oracle/Makefile."""
    subprocess.run(["make", "-C", str(_DIR), "_build/libmpr_oracle.so"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", str(_DIR), "ref"], check=True, capture_output=True)


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not ORACLE_SO.exists():
            build(ref=False)
        L = C.CDLL(str(ORACLE_SO))
        vp = C.c_void_p
        L.mpro_create.argtypes = [C.c_int, C.c_int64]
        L.mpro_create.restype = vp
        L.mpro_destroy.argtypes = [vp]
        L.mpro_render2d.argtypes = [vp, vp, C.c_int32, vp, C.c_float, C.c_int]
        L.mpro_render3d.argtypes = [vp, vp, C.c_int32, vp, C.c_int]
        L.mpro_render2d_brute.argtypes = [vp, vp, C.c_int32, vp, C.c_float, C.c_int]
        L.mpro_set_heat.argtypes = [vp, C.c_int]
        L.mpro_heat.argtypes = [vp]
        L.mpro_heat.restype = vp
        for name, rt in [("mpro_filled", vp), ("mpro_tiles", vp)]:
            getattr(L, name).argtypes = [vp, C.c_int]
            getattr(L, name).restype = rt
        L.mpro_tile_count.argtypes = [vp, C.c_int]
        L.mpro_tile_count.restype = C.c_uint64
        L.mpro_arena.argtypes = [vp]
        L.mpro_arena.restype = vp
        L.mpro_tape_index.argtypes = [vp]
        L.mpro_tape_index.restype = C.c_int32
        L.mpro_normals.argtypes = [vp]
        L.mpro_normals.restype = vp
        L.mpro_work_interval.argtypes = [vp]
        L.mpro_work_interval.restype = C.c_uint64
        L.mpro_work_float.argtypes = [vp]
        L.mpro_work_float.restype = C.c_uint64
        L.mpro_max_threads.restype = C.c_int
        L.mpro_tape_hash.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32)]
        L.mpro_tape_hash.restype = C.c_uint64
        L.mpro_tape_hashes.argtypes = [vp, vp, C.c_int64, vp, vp]
        L.mpro_tape_flatten.argtypes = [vp, C.c_int32, vp, C.c_int32]
        L.mpro_tape_flatten.restype = C.c_int32
        L.mpro_fx_ssao.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
        L.mpro_fx_shaded.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
        L.mpro_fx_ssao.restype = L.mpro_fx_shaded.restype = None
        _oracle = L
    return _oracle


def _as_array(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _mat(m, n):
    a = np.asarray(m, dtype=np.float32)
    assert a.shape == (n, n)
    return np.ascontiguousarray(a.T).reshape(-1)   # column-major


def view_matrix_3d():
    t = np.eye(4, dtype=np.float32)
    t[3, 2] = 0.3     # reference benchmark/render_3d_table.cpp:48-49
    return t


HEIGHTMAP_SO = _DIR / "_build" / "libheightmap_cpu.so"
_heightmap = None


def heightmap_cpu_ms(cells, dim: int, size: int, threads: int = 8, depth_out=None) -> float:
    """One frame of the libfive-algorithm CPU renderer (the stand-in the reference drivers link,
    oracle/heightmap_cabi.cpp) on `threads` host threads; returns its milliseconds."""
    global _heightmap
    if _heightmap is None:
        if not HEIGHTMAP_SO.exists():
            subprocess.run(["make", "-C", str(_DIR), "_build/libheightmap_cpu.so"], check=True, capture_output=True)
        L = C.CDLL(str(HEIGHTMAP_SO))
        L.mpro_heightmap_ms.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.mpro_heightmap_ms.restype = C.c_double
        _heightmap = L
    cells = np.ascontiguousarray(cells, dtype=np.uint64)
    ptr = depth_out.ctypes.data if depth_out is not None else None
    return float(_heightmap.mpro_heightmap_ms(cells.ctypes.data, cells.size, dim, size, threads, ptr))


def tape_hashes(arena: np.ndarray, starts: np.ndarray):
    """Hash + length of the logical tape starting at each arena index."""
    L = oracle_lib()
    starts = np.ascontiguousarray(starts, dtype=np.int32)
    hashes = np.zeros(starts.size, dtype=np.uint64)
    lens = np.zeros(starts.size, dtype=np.int32)
    if starts.size:
        L.mpro_tape_hashes(arena.ctypes.data, starts.ctypes.data, starts.size, hashes.ctypes.data, lens.ctypes.data)
    return hashes, lens


def tape_flatten(arena: np.ndarray, start: int) -> np.ndarray:
    L = oracle_lib()
    n = L.mpro_tape_flatten(arena.ctypes.data, int(start), None, 0)
    out = np.zeros(n, dtype=np.uint64)
    L.mpro_tape_flatten(arena.ctypes.data, int(start), out.ctypes.data, n)
    return out


class _Base:
    """Shared accessors; subclasses provide _filled/_tiles/... pointers."""
    size: int

    def image(self):
        return self.filled(3)

    def filled(self, stage):
        side = self.size // (64 >> (2 * stage))
        return _as_array(self._filled_ptr(stage), (side, side), np.int32)

    def tiles(self, stage):
        n = self._tile_count(stage)
        return _as_array(self._tiles_ptr(stage), (n,), TILE_DTYPE)

    def normals(self):
        return _as_array(self._normals_ptr(), (self.size, self.size), np.uint32)

    def arena(self):
        return _as_array(self._arena_ptr(), (self.tape_index(),), np.uint64)


class CpuOracle(_Base):
    """The C restatement, one frame at a time."""

    def __init__(self, size: int, num_subtapes: int = 0):
        self.L = oracle_lib()
        self.size = size
        self.h = self.L.mpro_create(size, num_subtapes)

    def close(self):
        if self.h:
            self.L.mpro_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render2D(self, cells, mat=None, z=0.0, threads=0):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        m = _mat(np.eye(3) if mat is None else mat, 3)
        self.L.mpro_render2d(self.h, cells.ctypes.data, cells.size, m.ctypes.data, z, threads)

    def render3D(self, cells, mat=None, threads=0):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        m = _mat(view_matrix_3d() if mat is None else mat, 4)
        self.L.mpro_render3d(self.h, cells.ctypes.data, cells.size, m.ctypes.data, threads)

    def render2D_brute(self, cells, mat=None, z=0.0, threads=0):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        m = _mat(np.eye(3) if mat is None else mat, 3)
        self.L.mpro_render2d_brute(self.h, cells.ctypes.data, cells.size, m.ctypes.data, z, threads)

    def _heatmap(self, render, cells, *args):
        """render*_heatmap (reference context.cu:1984-2340): amortised cells walked per pixel,
        divided by the clause count of the tape."""
        self.L.mpro_set_heat(self.h, 1)
        try:
            render(cells, *args)
            units = _as_array(self.L.mpro_heat(self.h), (self.size, self.size), np.uint64).copy()
        finally:
            self.L.mpro_set_heat(self.h, 0)
        return units

    def render2D_heatmap(self, cells, mat=None, z=0.0):
        units = self._heatmap(self.render2D, cells, mat, z)
        return heat_from_units(units, len(cells)), units

    def render3D_heatmap(self, cells, mat=None):
        units = self._heatmap(self.render3D, cells, mat)
        return heat_from_units(units, len(cells)), units

    def _filled_ptr(self, s): return self.L.mpro_filled(self.h, s)
    def _tiles_ptr(self, s): return self.L.mpro_tiles(self.h, s)
    def _tile_count(self, s): return int(self.L.mpro_tile_count(self.h, s))
    def _normals_ptr(self): return self.L.mpro_normals(self.h)
    def _arena_ptr(self): return self.L.mpro_arena(self.h)
    def tape_index(self): return int(self.L.mpro_tape_index(self.h))
    def work(self): return int(self.L.mpro_work_interval(self.h)), int(self.L.mpro_work_float(self.h))
    def max_threads(self): return int(self.L.mpro_max_threads())


def heat_from_units(units, n_cells):
    """Integer work units (1/4096 cell) -> the reference's float heatmap value."""
    return ((units.astype(np.float64) / 4096.0).astype(np.float32) / np.float32(n_cells - 2)).astype(np.float32)


def ref_available() -> bool:
    return REF_SO.exists()


def ref_lib():
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            raise RuntimeError(f"{REF_SO} not built (needs /root/reference; run `make -C oracle ref`)")
        L = C.CDLL(str(REF_SO))
        vp = C.c_void_p
        L.ref_ctx_create.argtypes = [C.c_int]
        L.ref_ctx_create.restype = vp
        L.ref_ctx_destroy.argtypes = [vp]
        L.ref_tape_create.argtypes = [vp, C.c_int32]
        L.ref_tape_create.restype = vp
        L.ref_tape_destroy.argtypes = [vp]
        L.ref_render2d.argtypes = [vp, vp, vp, C.c_float]
        L.ref_render3d.argtypes = [vp, vp, vp]
        L.ref_filled.argtypes = [vp, C.c_int]
        L.ref_filled.restype = vp
        L.ref_tiles.argtypes = [vp, C.c_int]
        L.ref_tiles.restype = vp
        L.ref_tile_array_size.argtypes = [vp, C.c_int]
        L.ref_tile_array_size.restype = C.c_uint64
        L.ref_tape_data.argtypes = [vp]
        L.ref_tape_data.restype = vp
        L.ref_tape_index.argtypes = [vp]
        L.ref_tape_index.restype = C.c_int32
        L.ref_num_active.argtypes = [vp]
        L.ref_num_active.restype = C.c_int32
        L.ref_normals.argtypes = [vp]
        L.ref_normals.restype = vp
        L.ref_num_subtapes.restype = C.c_int64
        L.ref_download.argtypes = [vp, vp, vp]
        L.ref_render2d_brute.argtypes = [vp, vp, vp, C.c_float]
        L.ref_render2d_heatmap.argtypes = [vp, vp, vp, C.c_float, vp]
        L.ref_render3d_heatmap.argtypes = [vp, vp, vp, vp]
        if hasattr(L, "ref_effects_create"):          # libraries built before effects.cu joined the recipe lack these
            L.ref_effects_create.restype = vp
            L.ref_effects_destroy.argtypes = [vp]
            L.ref_effects_samples.argtypes = [vp, vp, vp]
            L.ref_effects_draw_ssao.argtypes = [vp, vp]
            L.ref_effects_draw_shaded.argtypes = [vp, vp]
            L.ref_effects_image.argtypes = [vp]
            L.ref_effects_image.restype = vp
            L.ref_effects_tmp.argtypes = [vp]
            L.ref_effects_tmp.restype = vp
        _ref = L
    return _ref


class RefGpu(_Base):
    """The unmodified reference renderer (CUDA) behind oracle/ref_wrap.cu.  GPU box only."""

    def __init__(self, size: int):
        self.L = ref_lib()
        self.size = size
        self.h = self.L.ref_ctx_create(size)
        self._tapes = {}
        self._counts = [0, 0, 0, 0]

    def close(self):
        if self.h:
            for t in self._tapes.values():
                self.L.ref_tape_destroy(t[0])
            self._tapes = {}
            self.L.ref_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tape(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.uint64)
        key = (cells.ctypes.data, cells.size)
        if key not in self._tapes:
            self._tapes[key] = (self.L.ref_tape_create(cells.ctypes.data, cells.size), cells)
        return self._tapes[key][0]

    def render2D(self, cells, mat=None, z=0.0):
        m = _mat(np.eye(3) if mat is None else mat, 3)
        self.L.ref_render2d(self.h, self.tape(cells), m.ctypes.data, z)
        self._dim = 2

    def render3D(self, cells, mat=None):
        m = _mat(view_matrix_3d() if mat is None else mat, 4)
        self.L.ref_render3d(self.h, self.tape(cells), m.ctypes.data)
        self._dim = 3

    def render2D_brute(self, cells, mat=None, z=0.0):
        m = _mat(np.eye(3) if mat is None else mat, 3)
        self.L.ref_render2d_brute(self.h, self.tape(cells), m.ctypes.data, z)
        self._dim = 2

    def render2D_heatmap(self, cells, mat=None, z=0.0):
        m = _mat(np.eye(3) if mat is None else mat, 3)
        heat = np.zeros((self.size, self.size), dtype=np.float32)
        self.L.ref_render2d_heatmap(self.h, self.tape(cells), m.ctypes.data, z, heat.ctypes.data)
        self._dim = 2
        return heat

    def render3D_heatmap(self, cells, mat=None):
        m = _mat(view_matrix_3d() if mat is None else mat, 4)
        heat = np.zeros((self.size, self.size), dtype=np.float32)
        self.L.ref_render3d_heatmap(self.h, self.tape(cells), m.ctypes.data, heat.ctypes.data)
        self._dim = 3
        return heat

    def _filled_ptr(self, s): return self.L.ref_filled(self.h, s)
    def _tiles_ptr(self, s): return self.L.ref_tiles(self.h, s)

    def _tile_count(self, s):
        """Entries of stage s valid after the last frame.  The reference only keeps a
        high-water tile_array_size, so counts are rebuilt from the `next` chain."""
        tps0 = self.size // 64
        if s == 0:
            return tps0 ** self._dim
        chain = [0, 1, 2, 3] if self._dim == 3 else [0, 2, 3]
        if s not in chain:
            return 0
        prev = chain[chain.index(s) - 1]
        n_prev = self._tile_count(prev)
        t = _as_array(self._tiles_ptr(prev), (n_prev,), TILE_DTYPE)
        if s == 3:
            # survivors of the last interval level: their `next` was reset to -1 by
            # copy_active_tiles, so count positions instead
            return int((t["position"] != -1).sum())
        n_active = int(t["next"].max()) + 1 if n_prev else 0
        return n_active * 64

    def download(self, image_out, normals_out=None):
        self.L.ref_download(self.h, image_out.ctypes.data,
                            normals_out.ctypes.data if normals_out is not None else None)

    def drop_tapes(self):
        for t in self._tapes.values():
            self.L.ref_tape_destroy(t[0])
        self._tapes = {}

    def _normals_ptr(self): return self.L.ref_normals(self.h)
    def _arena_ptr(self): return self.L.ref_tape_data(self.h)
    def tape_index(self): return int(self.L.ref_tape_index(self.h))


class RefEffects:
    """The unmodified reference mpr::Effects (src/effects.cu compiled into oracle/_ref against the
    Eigen stand-in in oracle/shim).  GPU box only."""

    def __init__(self):
        self.L = ref_lib()
        if not hasattr(self.L, "ref_effects_create"):
            raise RuntimeError("oracle/_ref was built without src/effects.cu; run `make -C oracle ref`")
        self.h = self.L.ref_effects_create()

    def samples(self):
        """(kernel 64x3, rvecs 256x3) as Fortran-ordered float32 - the constructor's rand() draws."""
        k = np.zeros((64, 3), dtype=np.float32, order="F")
        r = np.zeros((256, 3), dtype=np.float32, order="F")
        self.L.ref_effects_samples(self.h, k.ctypes.data, r.ctypes.data)
        return k, r

    def draw(self, ctx: "RefGpu", shaded=False):
        (self.L.ref_effects_draw_shaded if shaded else self.L.ref_effects_draw_ssao)(self.h, ctx.h)
        s = ctx.size
        return (_as_array(self.L.ref_effects_image(self.h), (s, s), np.int32).copy(),
                _as_array(self.L.ref_effects_tmp(self.h), (s, s), np.int32).copy())

    def close(self):
        if self.h:
            self.L.ref_effects_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def effects(depth, normals, kernel, rvecs, shaded=False):
    """CPU restatement of mpr::Effects::drawSSAO / drawShaded (reference src/effects.cu:253-297) on a
    depth image + packed normals; kernel 64x3 and rvecs 256x3 as Fortran-ordered float32.  Returns
    (image, tmp)."""
    L = oracle_lib()
    depth = np.ascontiguousarray(depth, dtype=np.int32)
    normals = np.ascontiguousarray(normals, dtype=np.uint32)
    kernel = np.asfortranarray(kernel, dtype=np.float32)
    rvecs = np.asfortranarray(rvecs, dtype=np.float32)
    size = depth.shape[0]
    tmp = np.zeros((size, size), dtype=np.int32)
    image = np.zeros((size, size), dtype=np.int32)
    fn = L.mpro_fx_shaded if shaded else L.mpro_fx_ssao
    fn(depth.ctypes.data, normals.ctypes.data, kernel.ctypes.data, rvecs.ctypes.data, size,
       tmp.ctypes.data, image.ctypes.data)
    return image, tmp
