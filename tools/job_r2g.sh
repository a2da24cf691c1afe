set -x
O=gpurun_out/r2g; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or effects" 2>&1 | tail -12
for G in 1 2 4; do
  MPRB_FLOAT_GROUP=$G python tools/kernel_times.py bear:3:1024 hello_world:3:1024 architecture:3:2048 prospero:2:256 involute_gear_2d:2:1024 2>&1 | cut -c1-330 | tee -a $O/kt.log
done
