"""Frame-to-frame spread of small frames: gpu_ms (begin / end events of the frame) of N back-to-back frames,
and the per-kernel times of the fastest and the slowest one.  usage: python tools/frame_jitter.py model:dim:size ..."""
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

N = int(__import__("os").environ.get("JITTER_FRAMES", "64"))


def main():
    for case in sys.argv[1:]:
        model, dim, size = case.split(":")
        dim, size = int(dim), int(size)
        ctx = capi.Context(size, num_subtapes=6400000)
        tape = capi.Tape(parity.load_tape(model))
        render = (lambda: ctx.render2D(tape)) if dim == 2 else (lambda: ctx.render3D(tape))
        for _ in range(5):
            render()
        for timing in (False, True):
            ctx.set_timing(timing)
            gpu, wall, ks, sub = [], [], [], []
            for _ in range(N):
                t0 = time.perf_counter()
                render()
                wall.append((time.perf_counter() - t0) * 1e3)
                st = ctx.stats()
                gpu.append(st.gpu_ms)
                sub.append((int(st.i_sub_tiles), int(st.n_active[0]), int(st.p_written)))
                ks.append([round(float(v), 4) for v in list(st.kernel_ms)[: st.n_launches]])
            gpu = np.array(gpu)
            lo, hi = int(np.argmin(gpu)), int(np.argmax(gpu))
            step = max(1, N // 64)
            print(case, "timing" if timing else "plain", f"gpu_ms (every {step}th)", " ".join(f"{g:.2f}" for g in gpu[::step]), flush=True)
            print(case, "wall_ms", " ".join(f"{w:.2f}" for w in wall[::step]), flush=True)
            print(case, "first 96 frames: gpu_ms / tiles taken by k_eval_sub",
                  " ".join(f"{g:.2f}/{u[0]}" for g, u in zip(gpu[:96], sub[:96])), flush=True)
            print(case, "n_active[0], cells written (first, last frame)", sub[0][1:], sub[-1][1:], flush=True)
            print(case, "gpu_ms quartiles", np.percentile(gpu, [0, 25, 50, 75, 100]).round(3).tolist(), flush=True)
            if timing:
                print(case, "fastest", json.dumps(ks[lo]), "slowest", json.dumps(ks[hi]), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
