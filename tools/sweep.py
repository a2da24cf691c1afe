"""Frame-time sweep over the reference's table configurations, both arms, same process order.
Protocol: benchmark/stats.cpp (warm-up + timed frames, host clock around the render call incl.
its device sync).  Writes gpurun_out/sweep.md + sweep.json."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import oracle  # noqa: E402
import parity  # noqa: E402
from mpr_b200 import capi  # noqa: E402

SWEEP = [("prospero", 2, [256, 512, 1024, 2048, 3072, 4096]),
         ("involute_gear_2d", 2, [256, 512, 1024, 2048, 3072, 4096]),
         ("hello_world", 2, [1024, 4096]),
         ("bear", 3, [256, 512, 1024, 1536, 2048]),
         ("architecture", 3, [256, 512, 1024, 2048]),
         ("involute_gear_3d", 3, [256, 512, 1024, 2048]),
         ("hello_world", 3, [512, 1024])]
WARM, ITERS = 5, 30


def stats(fn):
    for _ in range(WARM):
        fn()
    ts = []
    for _ in range(ITERS):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.mean(ts)), float(np.std(ts, ddof=1)), float(np.min(ts))


def main():
    rows = []
    for model, dim, sizes in SWEEP:
        cells = parity.load_tape(model)
        for size in sizes:
            if size % 64:
                continue
            ref = oracle.RefGpu(size)
            f = (lambda: ref.render2D(cells)) if dim == 2 else (lambda: ref.render3D(cells))
            r = stats(f)
            ti_ref = ref.tape_index()
            ref.close()
            ctx = capi.Context(size, num_subtapes=6400000)
            tape = capi.Tape(cells)
            g = (lambda: ctx.render2D(tape)) if dim == 2 else (lambda: ctx.render3D(tape))
            m = stats(g)
            st = ctx.stats()
            rows.append(dict(model=model, dim=dim, size=size, ref_ms=r, mine_ms=m, speedup=r[0] / m[0],
                             gpu_ms=st.gpu_ms, arena_cells=st.tape_index, ref_arena_cells=ti_ref,
                             arena_full=bool(ti_ref >= 6400000 * 64 - 64)))
            print(rows[-1], flush=True)
            ctx.close()
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "sweep.json").write_text(json.dumps(rows, indent=1))
    with open(out / "sweep.md", "w") as f:
        f.write("| model | dim | size | reference CUDA build ms (σ) | mprb ms (σ) | speed-up |\n|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['model']} | {r['dim']}D | {r['size']} | {r['ref_ms'][0]:.3f} ({r['ref_ms'][1]:.3f}) | "
                    f"{r['mine_ms'][0]:.3f} ({r['mine_ms'][1]:.3f}) | {r['speedup']:.2f}x"
                    + (" (reference arena full)" if r["arena_full"] else "") + " |\n")


if __name__ == "__main__":
    main()
