"""Renders one case with libmprb and prints {"image": sha, "normals": sha, stats...} as JSON (helper for
tests that need a fresh process, e.g. to set MPRB_FLOAT_GROUP)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import parity  # noqa: E402
from mpr_b200 import capi  # noqa: E402


def main():
    model, dim, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    ctx = capi.Context(size, num_subtapes=6400000)
    tape = capi.Tape(parity.load_tape(model))
    (ctx.render2D if dim == 2 else ctx.render3D)(tape)
    st = ctx.stats()
    out = {"image": parity.digest(ctx.image()), "f_tiles": int(st.f_tiles), "f_items": int(st.f_items),
           "p_kept": int(sum(st.p_kept)), "p_written": int(st.p_written), "i_tiles": [int(v) for v in st.i_tiles],
           "i_sub_tiles": int(st.i_sub_tiles), "n_active": [int(v) for v in st.n_active]}
    if dim == 3:
        out["normals"] = parity.digest(ctx.normals())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
