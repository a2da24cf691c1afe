"""Per-kernel CUDA-event times of this repository's frames (no reference arm): quick A/B tool.
usage: python tools/kernel_times.py [model:dim:size ...]   -> gpurun_out/kernel_times.json"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import parity  # noqa: E402
from mpr_b200 import capi  # noqa: E402

DEFAULT = ["bear:3:1024", "bear:3:512", "hello_world:3:1024", "architecture:3:2048", "involute_gear_3d:3:1024",
           "prospero:2:4096", "involute_gear_2d:2:4096", "hello_world:2:4096"]
# intervals between the timer marks of api.cu:render(): the first one covers frame setup + the level-0 pass
NAMES3 = ["setup+L0 eval", "L0 rank", "L0 up", "L1 eval", "L1 rank", "L1 up", "L2 eval", "L2 rank", "L2 up",
          "float", "normals"]
NAMES2 = ["setup+L0 eval", "L0 rank", "L0 up", "L1 eval", "L1 rank", "L1 up", "float"]


def main():
    cases = sys.argv[1:] or DEFAULT
    out = {}
    for case in cases:
        model, dim, size = case.split(":")
        dim, size = int(dim), int(size)
        shard = {}
        if os.environ.get("MPRB_KT_SHARD"):                   # "world:rank": time one rank's share on one GPU
            from mpr_b200 import sharding
            world, rank = map(int, os.environ["MPRB_KT_SHARD"].split(":"))
            shard = sharding.diagonal_tiles(size, world, rank)
        ctx = capi.Context(size, num_subtapes=6400000, **shard)
        tape = capi.Tape(parity.load_tape(model))
        render = (lambda: ctx.render2D(tape)) if dim == 2 else (lambda: ctx.render3D(tape))
        for _ in range(5):
            render()
        ctx.set_timing(True)
        gpu, ks = [], []
        for _ in range(20):
            render()
            st = ctx.stats()
            gpu.append(st.gpu_ms)
            ks.append(list(st.kernel_ms)[: st.n_launches])
        k = np.median(np.array(ks), axis=0)
        names = NAMES3 if dim == 3 else NAMES2
        row = {"gpu_ms": float(np.median(gpu)), "gpu_ms_min": float(np.min(gpu)),
               "kernels": {n: round(float(v), 4) for n, v in zip(names, k)},
               "float_tiles": int(st.f_tiles), "float_cells": int(st.f_cells), "float_items": int(st.f_items),
               "push_tiles": list(st.p_tiles), "push_kept": list(st.p_kept), "push_written": int(st.p_written),
               "n_active": list(st.n_active), "i_sub_tiles": int(st.i_sub_tiles), "interval_tiles": list(st.i_tiles), "interval_cells": list(st.i_cells)}
        out[case] = row
        print(case, json.dumps(row), flush=True)
        ctx.close()
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "kernel_times.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
