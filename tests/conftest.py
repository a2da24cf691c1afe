import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tools"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"
MODELS = ["prospero", "involute_gear_2d", "involute_gear_3d", "architecture", "bear", "hello_world"]
# Models whose tapes contain no libdevice transcendentals: CPU restatement must be bit-exact.
EXACT_ON_CPU = {"prospero", "architecture", "hello_world"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_tape(model):
    return np.fromfile(GOLDEN / "tapes" / f"{model}.u64", dtype="<u8")


def golden_cases():
    """[(case, model, dim, size, summary_dict, npz_path_or_None)] minted from the reference build."""
    out = []
    for js in sorted((GOLDEN / "ref").glob("*.json")):
        case = js.stem
        model, dim, size = case.rsplit("_", 2)
        npz = js.with_suffix(".npz")
        out.append((case, model, int(dim[0]), int(size), json.loads(js.read_text()), npz if npz.exists() else None))
    return out


@pytest.fixture(scope="session")
def tapes():
    return {m: load_tape(m) for m in MODELS}
