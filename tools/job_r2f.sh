set -x
O=gpurun_out/r2f; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "fixture" 2>&1 | tail -3
for G in 1 2; do
  MPRB_FLOAT_GROUP=$G python tools/kernel_times.py bear:3:1024 hello_world:3:1024 architecture:3:2048 prospero:2:256 involute_gear_2d:2:1024 2>&1 | cut -c1-330 | tee -a $O/kt.log
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/run_one.py --model prospero --dim 2 --size 256 --frames 1 --subtapes 64000 2>&1 | grep -v "Saved host\|Host Frame\|^=========         " | tail -12
