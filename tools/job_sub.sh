set -x
O=gpurun_out/r2sub; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "clause_parallel" > $O/pytest_sub.log 2>&1; tail -15 $O/pytest_sub.log
python tools/kernel_times.py prospero:2:256 prospero:2:512 prospero:2:1024 hello_world:2:1024 bear:3:256 bear:3:1024 prospero:2:4096 2>&1 | cut -c1-900
MPRB_KT_SHARD=8:3 python tools/kernel_times.py bear:3:1024 prospero:2:4096 2>&1 | cut -c1-900
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
