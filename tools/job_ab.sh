set -x
O=gpurun_out/r2ab; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "clause_parallel or publish or work_items" > $O/pytest_sub.log 2>&1; tail -3 $O/pytest_sub.log
CASES="prospero:2:256 prospero:2:512 prospero:2:1024 hello_world:2:1024 bear:3:256 bear:3:1024 prospero:2:4096"
for L in build/old/libmprb.so mpr_b200/libmprb.so; do
echo "== LIB $L"
MPRB_LIBRARY=$PWD/$L python tools/kernel_times.py $CASES 2>&1 | cut -c1-420
MPRB_LIBRARY=$PWD/$L MPRB_KT_SHARD=8:3 python tools/kernel_times.py bear:3:1024 prospero:2:4096 2>&1 | cut -c1-420
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
