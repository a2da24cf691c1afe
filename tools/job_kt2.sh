set -x
python tools/kernel_times.py bear:3:256 bear:3:512 bear:3:1024 architecture:3:1024 hello_world:3:1024 involute_gear_3d:3:1024 2>&1 | cut -c1-420
MPRB_KT_SHARD=8:3 python tools/kernel_times.py bear:3:1024 2>&1 | cut -c1-420
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "clause_parallel" 2>&1 | tail -2
