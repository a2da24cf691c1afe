set -x
O=gpurun_out/r2ev; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "clause_parallel or fixture" 2>&1 | tail -2
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/run_one.py --model bear --dim 3 --size 128 --frames 1 --subtapes 64000 > $O/racecheck_bear_ptxloop.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/run_one.py --model prospero --dim 2 --size 256 --frames 1 --subtapes 64000 > $O/racecheck_prospero_remap_sub.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/run_one.py --model prospero --dim 2 --size 256 --frames 2 --subtapes 64000 > $O/memcheck_prospero_sub.log 2>&1
tail -2 $O/racecheck_bear_ptxloop.log $O/racecheck_prospero_remap_sub.log $O/memcheck_prospero_sub.log
grep "Error\|Warning" $O/racecheck_bear_ptxloop.log $O/racecheck_prospero_remap_sub.log | head
for G in 2 4; do MPRB_FLOAT_TMEM=0 MPRB_FLOAT_GROUP=$G MPRB_SUB_WAVES=0 python tools/frame_digest.py bear 3 1024; done
python tools/kernel_times.py prospero:2:256 prospero:2:512 bear:3:256 2>&1 | cut -c1-400
