"""The CPU restatement (oracle/mpr_oracle.c) against fixtures minted from the UNMODIFIED
reference CUDA build on a B200 (tests/golden/ref, written by tools/gpu_check.py through
oracle/_ref), plus known-answer tests for the interval operator table the reference defines
(reference inc/gpu_interval.hpp) and structural invariants of pushed tapes."""
import ctypes as C
import math
from fractions import Fraction

import numpy as np
import pytest

import oracle
import parity
from conftest import EXACT_ON_CPU, golden_cases, load_tape

SMALL = [c for c in golden_cases() if c[5] is not None]
OPS = dict(SQUARE=2, SQRT=3, NEG=4, SIN=5, COS=6, ASIN=7, ACOS=8, ATAN=9, EXP=10, ABS=11, LOG=12,
           ADD_LI=13, ADD_LR=14, MUL_LI=15, MUL_LR=16, MIN_LI=17, MIN_LR=18, MAX_LI=19, MAX_LR=20,
           SUB_LI=21, SUB_IR=22, SUB_LR=23, DIV_LI=24, DIV_IR=25, DIV_LR=26)


def iop(name, a, b=(0.0, 0.0)):
    L = oracle.oracle_lib()
    fa = (C.c_float * 2)(*a)
    fb = (C.c_float * 2)(*b)
    out = (C.c_float * 2)()
    L.mpro_interval_op.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    ch = L.mpro_interval_op(OPS[name], fa, fb, out)
    return (out[0], out[1]), ch


# ---- known answers -----------------------------------------------------------------------

def test_interval_sign_cases_of_multiplication():
    assert iop("MUL_LR", (2, 3), (4, 5))[0] == (8, 15)            # P * P
    assert iop("MUL_LR", (-3, -2), (4, 5))[0] == (-15, -8)        # N * P
    assert iop("MUL_LR", (-3, -2), (-5, -4))[0] == (8, 15)        # N * N
    assert iop("MUL_LR", (-2, 3), (4, 5))[0] == (-10, 15)         # M * P
    assert iop("MUL_LR", (-2, 3), (-5, -4))[0] == (-15, 10)       # M * N
    assert iop("MUL_LR", (-2, 3), (-5, 7))[0] == (-15, 21)        # M * M: min/max of two candidates
    assert iop("MUL_LR", (0, 0), (-5, 7))[0] == (0, 0)            # zero-ish operand
    assert iop("MUL_LR", (2, 3), (0, 0))[0] == (0, 0)
    nan = float("nan")
    assert iop("MUL_LR", (nan, nan), (1, 2))[0] == (0, 0)         # NaN fails every strict test
    assert iop("MUL_LI", (-2, 3), (-2, 0))[0] == (-6, 4)
    assert iop("MUL_LI", (-2, 3), (2, 0))[0] == (-4, 6)


def test_interval_division_and_poles():
    inf = float("inf")
    assert iop("DIV_LR", (1, 2), (-1, 1))[0] == (-inf, inf)
    assert iop("DIV_LR", (1, 2), (0, 1))[0] == (-inf, inf)        # divisor touching zero
    assert iop("DIV_LR", (2, 4), (1, 2))[0] == (1, 4)
    assert iop("DIV_LR", (-4, -2), (1, 2))[0] == (-4, -1)
    assert iop("DIV_LR", (-4, 2), (-2, -1))[0] == (-2, 4)
    assert iop("DIV_LI", (2, 4), (0, 0))[0] == (-inf, inf)
    assert iop("DIV_LI", (2, 4), (-2, 0))[0] == (-2, -1)
    assert iop("DIV_IR", (2, 4), (8, 0))[0] == (2, 4)             # imm / x


def test_interval_min_max_report_which_side_won():
    assert iop("MIN_LR", (0, 1), (2, 3)) == ((0, 1), 1)
    assert iop("MIN_LR", (2, 3), (0, 1)) == ((0, 1), 2)
    assert iop("MIN_LR", (0, 2), (1, 3)) == ((0, 2), 0)
    assert iop("MIN_LR", (0, 1), (1, 3))[1] == 0                  # touching is not strict: undecided
    assert iop("MAX_LR", (2, 3), (0, 1)) == ((2, 3), 1)
    assert iop("MAX_LR", (0, 1), (2, 3)) == ((2, 3), 2)
    assert iop("MAX_LR", (0, 2), (1, 3)) == ((1, 3), 0)
    assert iop("MIN_LI", (0, 1), (5, 0)) == ((0, 1), 1)
    assert iop("MIN_LI", (6, 7), (5, 0)) == ((5, 5), 2)
    assert iop("MAX_LI", (0, 1), (5, 0)) == ((5, 5), 2)


def test_interval_unary_quirks_of_the_reference():
    nan = float("nan")
    assert iop("COS", (0.1, 0.2))[0] == (-1, 1)                   # gpu_interval.hpp:353 early return
    assert iop("SIN", (0.1, 0.2))[0] == (-1, 1)
    lo, hi = iop("LOG", (-1, 1))[0]
    assert lo == 0 and hi == 0                                    # lower bound clamps to 0, not -inf
    assert all(math.isnan(v) for v in iop("LOG", (-2, -1))[0])
    assert all(math.isnan(v) for v in iop("SQRT", (-2, -1))[0])
    assert iop("SQRT", (-1, 4))[0] == (0, 2)
    assert iop("SQUARE", (-3, 2))[0] == (0, 9)
    assert iop("SQUARE", (-2, 3))[0] == (0, 9)
    assert iop("SQUARE", (-3, -2))[0] == (4, 9)
    assert iop("ABS", (-3, 2))[0] == (0, 3)
    assert iop("ABS", (-3, -2))[0] == (2, 3)
    assert iop("NEG", (-3, 2))[0] == (-2, 3)
    assert all(math.isnan(v) for v in iop("ACOS", (1.5, 2))[0])
    lo, hi = iop("ACOS", (-1, 1))[0]
    assert lo == 0 and abs(hi - math.pi) < 1e-6                   # decreasing: bounds swap
    assert iop("SUB_IR", (1, 2), (5, 0))[0] == (3, 4)


@pytest.mark.parametrize("name", ["ADD_LR", "SUB_LR", "MUL_LR", "DIV_LR"])
def test_directed_rounding_encloses_the_exact_result(name):
    rng = np.random.default_rng(7)
    f = {"ADD_LR": lambda x, y: x + y, "SUB_LR": lambda x, y: x - y, "MUL_LR": lambda x, y: x * y,
         "DIV_LR": lambda x, y: x / y}[name]
    for _ in range(300):
        a = np.sort(rng.normal(size=2).astype(np.float32) * 10)
        b = np.sort(rng.normal(size=2).astype(np.float32) * 10)
        if name == "DIV_LR" and b[0] <= 0 <= b[1]:
            continue
        (lo, hi), _ = iop(name, tuple(map(float, a)), tuple(map(float, b)))
        exact = [f(Fraction(float(x)), Fraction(float(y))) for x in a for y in b]
        assert Fraction(lo) <= min(exact) and max(exact) <= Fraction(hi)
        # and tight: at most one float away from the exact bound
        assert np.nextafter(np.float32(lo), np.float32(np.inf)) >= np.float32(float(min(exact)))
        assert np.nextafter(np.float32(hi), np.float32(-np.inf)) <= np.float32(float(max(exact)))


# ---- against the reference build ----------------------------------------------------------

@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
def test_cpu_restatement_matches_reference_fixture(case):
    name, model, dim, size, summary, npz = case
    ref = dict(np.load(npz))
    ref["dim"], ref["size"] = dim, size
    o = oracle.CpuOracle(size, 6400000)
    (o.render2D if dim == 2 else o.render3D)(load_tape(model))
    fp = parity.fingerprint(o, dim)
    bad = parity.compare(ref, fp, normals_lsb=0 if model in EXACT_ON_CPU else 1)
    if model in EXACT_ON_CPU:
        assert not bad, bad
    else:
        # libdevice vs glibc transcendentals may flip isolated pixels; everything that does
        # not depend on them must still agree, and the image in all but a handful of pixels
        for k, v in bad.items():
            assert k in ("image", "normals"), (k, v)
        diff = int((ref["image"] != fp["image"]).sum())
        assert diff <= max(4, ref["image"].size // 10000), diff
    assert parity.summarize(ref)["image"] == summary["image"]
    o.close()


def test_frames_are_reproducible_and_thread_count_independent():
    cells = load_tape("hello_world")
    a = oracle.CpuOracle(256)
    b = oracle.CpuOracle(256)
    a.render3D(cells, threads=1)
    b.render3D(cells, threads=4)
    assert not parity.compare(parity.fingerprint(a, 3), parity.fingerprint(b, 3))
    b.render3D(cells, threads=4)      # same context again
    assert not parity.compare(parity.fingerprint(a, 3), parity.fingerprint(b, 3))


# ---- structural invariants of tape shortening -----------------------

@pytest.mark.parametrize("model,dim,size", [("prospero", 2, 256), ("architecture", 3, 128)])
def test_pushed_tapes_are_rewritten_subsequences_of_their_parent(model, dim, size):
    cells = load_tape(model)
    o = oracle.CpuOracle(size)
    (o.render2D if dim == 2 else o.render3D)(cells)
    arena = np.ascontiguousarray(o.arena())
    root = [int(c) for c in cells[1:-1]]
    stage = 0
    tiles = o.tiles(stage)
    active = tiles[tiles["position"] != -1]
    assert len(active)
    for t in active[:40]:
        flat = oracle.tape_flatten(arena, int(t["tape"]))
        assert int(flat[0]) == int(cells[0]) and (int(flat[-1]) & 0xFF) == 0
        body = [int(c) for c in flat[1:-1]]
        assert len(body) <= len(root)
        it = iter(root)
        for c in body:
            op = c & 0xFF
            assert op != 1                                     # flattening removed every JUMP
            for r in it:                                       # same order as in the parent
                same_fields = (r >> 8) == (c >> 8)
                if same_fields and (op == (r & 0xFF) or (op in (27, 28, 29) and 17 <= (r & 0xFF) <= 20)):
                    break
            else:
                pytest.fail("pushed clause is not an (op-rewritten) clause of the parent, in order")
    o.close()


def test_tape_evaluates_the_expression_it_was_built_from():
    from test_host import _frep
    from mpr_b200 import capi
    X, Y, SQUARE, SQRT, ADD, MAX, SUB = 2, 3, 7, 8, 20, 23, 24
    nodes = [(X,), (Y,), (SQUARE, 0), (SQUARE, 1), (ADD, 2, 3), (SQRT, 4), ("const", 1.0), ("const", 0.5),
             (SUB, 5, 6), (SUB, 7, 5), (MAX, 8, 9)]
    cells = capi.tape_from_frep(_frep(nodes))
    rng = np.random.default_rng(3)
    pts = rng.uniform(-2, 2, size=(1000, 3)).astype(np.float32)
    out = np.zeros(1000, dtype=np.float32)
    L = oracle.oracle_lib()
    L.mpro_eval_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.mpro_eval_points(cells.ctypes.data, pts.ctypes.data, 1000, out.ctypes.data)
    r = np.sqrt(pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1])
    want = np.maximum(r - np.float32(1.0), np.float32(0.5) - r)
    assert np.array_equal(out, want.astype(np.float32))


# ---- post-effects restatement (reference src/effects.cu); parity unpinned, so these pin properties ----

def _plane(size, h, nz_only=True):
    depth = np.full((size, size), h, dtype=np.int32)
    normals = np.full((size, size), 128 | (128 << 8) | (255 << 16) | (0xff << 24), dtype=np.uint32)
    return depth, normals


def test_effect_samples_follow_the_reference_constructor():
    from mpr_b200.capi import glibc_effect_samples
    kernel, rvecs = glibc_effect_samples()
    assert kernel.shape == (64, 3) and rvecs.shape == (256, 3)
    i = np.arange(64, dtype=np.float32) / np.float32(63)
    np.testing.assert_allclose(np.linalg.norm(kernel, axis=1), i * i * 0.9 + 0.1, rtol=1e-5)   # effects.cu:236-238
    assert (kernel[:, 2] >= 0).all()                                                         # hemisphere
    np.testing.assert_allclose(np.linalg.norm(rvecs, axis=1), 1.0, rtol=1e-5)
    assert (rvecs[:, 2] == 0).all()
    k2, r2 = glibc_effect_samples()
    assert (k2 == kernel).all() and (r2 == rvecs).all()


def test_ssao_of_an_unoccluded_plane_is_white_and_of_nothing_is_black():
    from mpr_b200.capi import glibc_effect_samples
    kernel, rvecs = glibc_effect_samples()
    size = 64
    depth, normals = _plane(size, 20)
    image, tmp = oracle.effects(depth, normals, kernel, rvecs)
    # every sample lies above a plane facing +z (kernel z > 0): no occlusion, away from the frame edge
    assert (tmp[8:-8, 8:-8] == 255).all() and (image[8:-8, 8:-8] == 255).all()
    image, tmp = oracle.effects(np.zeros_like(depth), normals, kernel, rvecs)
    assert not tmp.any() and not image.any()


def test_ssao_darkens_the_foot_of_a_step():
    from mpr_b200.capi import glibc_effect_samples
    kernel, rvecs = glibc_effect_samples()
    size = 128
    depth, normals = _plane(size, 20)
    depth[:, 64:] = 30                      # a wall 10 voxels high (0.16 in NDC, RADIUS = 0.1) right of x = 64
    image, tmp = oracle.effects(depth, normals, kernel, rvecs)
    near, far = tmp[32:96, 60:64].mean(), tmp[32:96, 20:40].mean()
    assert far == 255 and near < 240
    assert image[32:96, 61].mean() < 250    # the blur keeps the darkening


def test_shading_of_a_plane_matches_the_closed_form():
    from mpr_b200.capi import glibc_effect_samples
    kernel, rvecs = glibc_effect_samples()
    size = 64
    depth, normals = _plane(size, 20)
    image, _ = oracle.effects(depth, normals, kernel, rvecs, shaded=True)
    assert ((image >> 24) & 0xff == 0xff).all()
    grey = image & 0xff
    assert ((image >> 8) & 0xff == grey).all() and ((image >> 16) & 0xff == grey).all()
    xs = 2.0 * ((np.arange(size) + 0.5) / size - 0.5)
    px, py = np.meshgrid(xs, xs)
    pz = 2.0 * ((20 + 0.5) / size - 0.5)
    l = np.stack([5 - px, 5 - py, 10 - pz + 0 * px])
    want = np.clip(l[2] / np.linalg.norm(l, axis=0) * (127 / 127.0) * 0.8 + 0.2, 0, 1) * 255     # effects.cu:199-217
    inner = (slice(8, -8), slice(8, -8))
    assert np.abs(grey[inner] - np.floor(want[inner])).max() <= 1


# ---- analysis variants: brute force and the work meter (reference context.cu:1461-2340) ----

@pytest.mark.parametrize("model,size", [("hello_world", 128), ("prospero", 256)])
def test_brute_force_frame_equals_the_subdivided_one(model, size):
    cells = load_tape(model)
    o = oracle.CpuOracle(size, 100000)
    o.render2D(cells)
    fancy = o.image().copy()
    o.render2D_brute(cells)
    assert np.array_equal(o.image(), fancy)
    assert o._tile_count(3) == (size // 8) ** 2
    o.close()


def test_work_meter_charges_every_pixel_for_every_level_that_covers_it():
    cells = load_tape("prospero")
    n = len(cells) - 2
    size = 1024
    o = oracle.CpuOracle(size, 100000)
    o.render2D(cells)
    image = o.image().copy()
    heat, units = o.render2D_heatmap(cells)
    assert np.array_equal(o.image(), image)                 # metering does not change the frame
    # a tile spreads cells / px^2 over px^2 pixels, so the total is a whole number of cells
    assert int(units.sum()) % 4096 == 0
    # every pixel is covered by one level-0 tile, whose forward walk visits the whole tape:
    # n cells over 64 x 64 pixels
    assert units.min() >= n and heat.min() >= np.float32(1.0 / 4096) * np.float32(0.999)
    assert np.allclose(heat, units / 4096.0 / n, rtol=1e-6)
    # a level-0 tile that is proven empty or filled costs exactly one walk of the root tape
    t0 = o.tiles(0)
    dead = t0["position"] == -1
    assert dead.any()
    k = int(np.flatnonzero(dead)[0])
    ty, tx = divmod(k, size // 64)
    assert (units[ty * 64:(ty + 1) * 64, tx * 64:(tx + 1) * 64] == n).all()
    # ambiguous tiles were walked forwards and backwards, then their children were charged too
    k = int(np.flatnonzero(~dead)[0])
    ty, tx = divmod(k, size // 64)
    assert (units[ty * 64:(ty + 1) * 64, tx * 64:(tx + 1) * 64] >= 2 * n).all()
    o.close()
