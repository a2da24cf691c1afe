"""Host-side logic: the C-ABI library loads and exports what include/mprb.h declares, the
.frep reader + tape packer reproduce the committed tapes, argument validation."""
import ctypes as C
import hashlib
import json
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, MODELS, ROOT, load_tape
from mpr_b200 import capi

REF_FILES = Path("/root/reference/benchmark/files")


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "mprb.h").read_text()
    declared = set(re.findall(r"\b(mprb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in mprb.h but not exported"
    assert declared == set(capi.EXPORTS)
    assert b"sm_100a" in L.mprb_version()


def test_tape_fixture_index_matches_files():
    index = json.loads((GOLDEN / "tapes" / "index.json").read_text())
    for m in MODELS:
        cells = load_tape(m)
        assert cells.size == index[m]["cells"]
        assert hashlib.sha256(cells.tobytes()).hexdigest() == index[m]["sha256"]
        ops = (cells & 0xFF).astype(int)
        assert ops[0] == 0 and ops[-1] == 0                      # header / end cell
        assert ((ops[1:-1] >= 2) & (ops[1:-1] <= 26)).all()       # only clause opcodes, no COPY_*/JUMP
        assert int(((ops >= 17) & (ops <= 20)).sum()) == index[m]["choice_clauses"]


# Clause / slot / choice counts measured in the survey of the reference
SURVEY = {"prospero": (6056, 85, 2354), "involute_gear_2d": (1660, 67, 374), "involute_gear_3d": (1735, 84, 374),
          "architecture": (1296, 93, 488), "bear": (544, 23, 27), "hello_world": (328, 27, 97)}


@pytest.mark.skipif(not REF_FILES.exists(), reason="reference fixtures not present on this box")
@pytest.mark.parametrize("model", MODELS)
def test_frep_reader_and_packer_reproduce_committed_tapes(model):
    data = (REF_FILES / f"{model}.frep").read_bytes()
    cells = capi.tape_from_frep(data, simplify=True)
    assert np.array_equal(cells, load_tape(model))
    clauses, slots, choices = SURVEY[model]
    assert cells.size - 2 == clauses
    used = {int(b) for sh in (8, 16, 24) for b in ((cells >> sh) & 0xFF)}
    assert max(used) + 1 == slots
    ops = (cells & 0xFF).astype(int)
    assert int(((ops >= 17) & (ops <= 20)).sum()) == choices
    # the fixtures are stored already simplified: re-simplifying on load is a no-op
    assert np.array_equal(capi.tape_from_frep(data, simplify=False), cells)


def test_frep_reader_rejects_garbage():
    with pytest.raises(capi.MprbError):
        capi.tape_from_frep(b"\x00\x01\x02", simplify=True)


def _frep(nodes):
    """Builds a minimal archive: 'T' "" "" nodes 0xFF 0xFF (libfive deserializer.cpp:38-143)."""
    import struct
    out = b'T""""'
    for n in nodes:
        if n[0] == "const":
            out += bytes([1]) + struct.pack("<f", n[1])
        elif len(n) == 1:
            out += bytes([n[0]])
        elif len(n) == 2:
            out += bytes([n[0]]) + struct.pack("<I", n[1])
        else:   # binary: rhs is stored first
            out += bytes([n[0]]) + struct.pack("<II", n[2], n[1])
    return out + b"\xff\xff"


def test_packer_known_answer_circle():
    # max(sqrt(x^2 + y^2) - 1, 0.5 - sqrt(x^2 + y^2)): the default shape of the reference's
    # print_tape_table driver (benchmark/print_tape_table.cpp:29)
    X, Y, CONST, SQUARE, SQRT, ADD, MAX, SUB = 2, 3, 1, 7, 8, 20, 23, 24
    nodes = [(X,), (Y,), (SQUARE, 0), (SQUARE, 1), (ADD, 2, 3), (SQRT, 4), ("const", 1.0), ("const", 0.5),
             (SUB, 5, 6), (SUB, 7, 5), (MAX, 8, 9)]
    cells = capi.tape_from_frep(_frep(nodes))
    ops = [int(c & 0xFF) for c in cells]
    # header, square, square, add, sqrt, sub_lhs_imm, sub_imm_rhs (either order), max_lhs_rhs, end
    assert ops[0] == 0 and ops[-1] == 0 and len(ops) == 9
    assert sorted(ops[1:3]) == [2, 2] and ops[3] == 14 and ops[4] == 3
    assert sorted(ops[5:7]) == [21, 22] and ops[7] == 20
    hdr = int(cells[0])
    assert (hdr >> 8) & 0xFF and (hdr >> 16) & 0xFF and (hdr >> 24) == 0     # x, y bound; z unused
    imms = {int(c & 0xFF): np.frombuffer(np.uint32(int(c) >> 32).tobytes(), dtype="<f4")[0] for c in cells[5:7]}
    assert imms[21] == 1.0 and imms[22] == 0.5
    assert ((int(cells[-1]) >> 8) & 0xFF) == ((int(cells[7]) >> 8) & 0xFF)   # end cell names the result slot
    # operands whose last use is a clause are released first, so outputs reuse slots: <= 3 live + slot 0
    assert max(int((c >> 8) & 0xFF) for c in cells[1:-1]) <= 3


def test_packer_folds_constants_and_dedups():
    X, CONST, ADD, MUL = 2, 1, 20, 21
    # (x + (2 * 3)) * (x + 6): the product of two identical sub-expressions becomes a square
    nodes = [(X,), ("const", 2.0), ("const", 3.0), (MUL, 1, 2), (ADD, 0, 3), ("const", 6.0), (ADD, 0, 5), (MUL, 4, 6)]
    cells = capi.tape_from_frep(_frep(nodes))
    ops = [int(c & 0xFF) for c in cells[1:-1]]
    assert ops == [13, 2]      # ADD_LHS_IMM 6, SQUARE


def test_tape_create_validates_before_touching_the_device():
    L = capi.lib()
    h = C.c_void_p()
    bad = np.array([0, 99, 0], dtype=np.uint64)            # opcode 99 is not a clause
    assert L.mprb_tape_create(bad.ctypes.data, 3, C.byref(h)) == 2
    assert b"opcode" in L.mprb_last_error()
    short = np.array([0], dtype=np.uint64)
    assert L.mprb_tape_create(short.ctypes.data, 1, C.byref(h)) == 2
    no_end = np.array([0, 2 | (1 << 8) | (1 << 16), 5], dtype=np.uint64)
    assert L.mprb_tape_create(no_end.ctypes.data, 3, C.byref(h)) == 2


def test_ctx_create_rejects_bad_sizes():
    L = capi.lib()
    h = C.c_void_p()
    assert L.mprb_ctx_create(100, None, C.byref(h)) == 2
    assert L.mprb_ctx_create(0, None, C.byref(h)) == 2
    assert b"multiple of 64" in L.mprb_last_error()
