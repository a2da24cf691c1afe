"""One mprb context over 1, 2, 4, 8 GPUs of this process (mprb_ctx_opts::n_gpus): wall time per frame of
the synchronous render call (what a caller of mpr::Context sees) and the frame checked against the
single-GPU one.  Usage: python tools/multi_gpu_times.py bear:3:1024 prospero:2:4096"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import parity  # noqa: E402
from mpr_b200 import capi  # noqa: E402


def main():
    import torch
    n_dev = torch.cuda.device_count()
    out = {}
    for spec in sys.argv[1:] or ["bear:3:1024", "prospero:2:4096"]:
        model, dim, size = spec.split(":")
        dim, size = int(dim), int(size)
        cells = parity.load_tape(model)
        tape = capi.Tape(cells)
        want = None
        row = {}
        for n in (1, 2, 4, 8):
            if n > n_dev:
                break
            ctx = capi.Context(size, num_subtapes=6400000, n_gpus=n)
            render = ctx.render2D if dim == 2 else ctx.render3D
            for _ in range(5):
                render(tape)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter()
                render(tape)
                ts.append((time.perf_counter() - t0) * 1e3)
            img = np.array(ctx.image(), copy=True)
            nrm = np.array(ctx.normals(), copy=True) if dim == 3 else None
            if want is None:
                want = (img, nrm)
            same = bool(np.array_equal(img, want[0]) and (dim == 2 or np.array_equal(nrm, want[1])))
            row[n] = {"wall_ms_median": round(float(np.median(ts)), 4), "wall_ms_min": round(min(ts), 4),
                      "equals_single_gpu_frame": same}
            ctx.close()
        out[spec] = row
        print(spec, json.dumps(row), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/multi_gpu_times.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
