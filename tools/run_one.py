"""Renders one case a few times with either arm (for ncu / sanitizer runs)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import parity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="mine", choices=["mine", "ref"])
    ap.add_argument("--model", default="bear")
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--subtapes", type=int, default=6400000)
    a = ap.parse_args()
    cells = parity.load_tape(a.model)
    if a.impl == "ref":
        import oracle
        r = oracle.RefGpu(a.size)
        for _ in range(a.frames):
            (r.render2D if a.dim == 2 else r.render3D)(cells)
        print("ref tape_index", r.tape_index())
    else:
        from mpr_b200 import capi
        ctx = capi.Context(a.size, num_subtapes=a.subtapes)
        tape = capi.Tape(cells)
        for _ in range(a.frames):
            (ctx.render2D if a.dim == 2 else ctx.render3D)(tape)
        print("mine", ctx.stats().asdict())


if __name__ == "__main__":
    main()
