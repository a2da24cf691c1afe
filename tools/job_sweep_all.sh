# All six models at the reference's table sizes, both arms (tools/sweep.py), and per-kernel times of the
# models with renamed slots.  Outputs under gpurun_out/ (sweep.md / sweep.json / kernel_times.json).
set -x
export MPRB_LIBRARY=$PWD/build/old/libmprb.so
python tools/sweep.py 2>&1 | cut -c1-300 | tail -40
python tools/kernel_times.py architecture:3:2048 architecture:3:1024 involute_gear_3d:3:2048 involute_gear_3d:3:1024 involute_gear_2d:2:1024 involute_gear_2d:2:4096 hello_world:2:1024 2>&1 | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture" 2>&1 | tail -2
