set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture" 2>&1 | tail -2
CASES="architecture:3:2048 involute_gear_3d:3:2048 involute_gear_3d:3:1024 prospero:2:4096 involute_gear_2d:2:4096"
echo "== old"; MPRB_LIBRARY=$PWD/build/old/libmprb.so python tools/kernel_times.py $CASES 2>&1 | grep -o '^[a-z_0-9:]* \|"gpu_ms": [0-9.]*\|"float": [0-9.]*\|"normals": [0-9.]*' | paste -sd' ' | sed 's/ \([a-z_]*[0-9a-z_]*:[23]:\)/\n\1/g'
for R in 16 24 32 40 48 64; do
echo "== rows $R"; MPRB_FLOAT_ROWS=$R python tools/kernel_times.py $CASES 2>&1 | grep -o '^[a-z_0-9:]* \|"gpu_ms": [0-9.]*\|"float": [0-9.]*' | paste -sd' ' | sed 's/ \([a-z_]*[0-9a-z_]*:[23]:\)/\n\1/g'
done
