set -x
O=gpurun_out/r2e; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or live or cpu_restatement" 2>&1 | tail -4
for G in 1 2 4; do
  MPRB_FLOAT_GROUP=$G python tools/kernel_times.py bear:3:1024 hello_world:3:1024 architecture:3:2048 prospero:2:4096 prospero:2:256 involute_gear_2d:2:1024 2>&1 | cut -c1-460 | tee -a $O/kt.log
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/run_one.py --model hello_world --dim 3 --size 128 --frames 1 --subtapes 64000 2>&1 | tail -4
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/run_one.py --model prospero --dim 2 --size 256 --frames 1 --subtapes 64000 2>&1 | tail -4
