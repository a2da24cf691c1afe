# A/B of two builds of the library (build/old = before, mpr_b200 = after) + the frame-jitter probe.
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or work_items" 2>&1 | tail -2
for L in build/old/libmprb.so mpr_b200/libmprb.so; do
echo "== LIB $L"
MPRB_LIBRARY=$PWD/$L python tools/kernel_times.py bear:3:1024 bear:3:256 hello_world:3:1024 hello_world:2:1024 2>&1 | cut -c1-420
done
python tools/frame_jitter.py prospero:2:1024 hello_world:3:512 2>&1 | cut -c1-900
