set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or work_items or live_reference or clause_parallel" 2>&1 | tail -3
echo "== G4 hybrid (default)"
python tools/kernel_times.py bear:3:1024 hello_world:3:1024 bear:3:512 bear:3:256 prospero:2:256 hello_world:2:4096 2>&1 | cut -c1-420
echo "== G2 hybrid"
MPRB_FLOAT_TMEM_GROUP=2 python tools/kernel_times.py bear:3:1024 hello_world:3:1024 bear:3:512 hello_world:2:4096 2>&1 | cut -c1-420
