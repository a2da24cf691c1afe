set -x
O=gpurun_out/r2sub; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "clause_parallel or publish or work_items" > $O/pytest_sub.log 2>&1; tail -3 $O/pytest_sub.log
for W in 0 4; do
echo "== MPRB_SUB_WAVES=$W"
MPRB_SUB_WAVES=$W python tools/kernel_times.py prospero:2:256 prospero:2:512 hello_world:2:1024 bear:3:256 prospero:2:4096 2>&1 | cut -c1-420
MPRB_SUB_WAVES=$W MPRB_KT_SHARD=8:3 python tools/kernel_times.py bear:3:1024 prospero:2:4096 2>&1 | cut -c1-420
done
