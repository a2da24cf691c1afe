set -x
O=gpurun_out/r2c; mkdir -p $O
for U in 1 2; do for G in 1 2 4; do
  echo "=== G=$G U=$U" >> $O/sweep.log
  MPRB_FLOAT_GROUP=$G MPRB_FLOAT_UNROLL=$U python -m pytest tests/test_gpu_parity.py -x -q -k "fixture and (bear_3d_256 or hello_world or bear_3d_128)" 2>&1 | tail -2 >> $O/sweep.log
  MPRB_FLOAT_GROUP=$G MPRB_FLOAT_UNROLL=$U python tools/kernel_times.py bear:3:1024 hello_world:3:1024 hello_world:2:4096 2>&1 | cut -c1-420 >> $O/sweep.log
done; done
cat $O/sweep.log
