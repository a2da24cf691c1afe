set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or clause_parallel" 2>&1 | tail -2
for L in build/old/libmprb.so mpr_b200/libmprb.so; do
echo "== LIB $L"
MPRB_LIBRARY=$PWD/$L python tools/kernel_times.py prospero:2:4096 prospero:2:2048 bear:3:1024 architecture:3:1024 prospero:2:256 2>&1 | cut -c1-330
done
