"""Drop-in check on the GPU box: the reference's own benchmark drivers, compiled unchanged
against mpr_b200/inc and linked with libmprb.so (`make drivers`, prebuilt binaries travel in
build/drivers), run and produce their tables / images."""
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
DRV = ROOT / "build" / "drivers"


def run(name, cwd, *args, timeout=300):
    exe = DRV / name
    if not exe.exists():
        pytest.skip(f"{exe} not built (needs the reference sources at build time)")
    return subprocess.run([str(exe), *args], cwd=cwd, capture_output=True, text=True, timeout=timeout)


def test_render_2d_table_driver(tmp_path):
    r = run("render_2d_table", tmp_path)            # default shape: two spheres (render_2d_table.cpp:40-46)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in r.stdout.strip().splitlines()]
    assert [int(x[0]) for x in rows] == [256, 512, 1024, 2048, 3072, 4096]     # "size mean_ms stdev_ms"
    assert all(float(x[1]) > 0 for x in rows)
    png = (tmp_path / "out_gpu_256.png").read_bytes()
    assert png[:8] == b"\x89PNG\r\n\x1a\n" and len(png) > 256 * 256


def test_render_3d_table_driver(tmp_path):
    r = run("render_3d_table", tmp_path)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in r.stdout.strip().splitlines()]
    assert [int(x[0]) for x in rows][:3] == [256, 512, 1024]
    assert (tmp_path / "out_gpu_depth_256.png").exists() or any(tmp_path.glob("*.png"))


def test_print_tape_table_driver(tmp_path):
    r = run("print_tape_table", tmp_path)
    assert r.returncode == 0, r.stderr
    # LaTeX rows "op & out & lhs & rhs" for max(sqrt(x^2+y^2) - 1, 0.5 - sqrt(x^2+y^2))
    assert "SQRT" in r.stdout and "MAX" in r.stdout and "SUB" in r.stdout
