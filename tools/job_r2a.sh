set -x
mkdir -p gpurun_out/r2a
python tools/gpu_check.py --no-cpu --iters 5 --out gpurun_out/r2a --cases architecture_3d_2048,involute_gear_3d_3d_2048,involute_gear_2d_2d_256,involute_gear_2d_2d_1024,involute_gear_2d_2d_2048,involute_gear_2d_2d_3072,involute_gear_2d_2d_4096,bear_3d_1536,prospero_2d_2048,hello_world_2d_1024 > gpurun_out/r2a/check.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/run_one.py --model hello_world --dim 3 --size 128 --frames 1 --subtapes 64000 > gpurun_out/r2a/racecheck_hello3d.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/run_one.py --model prospero --dim 2 --size 256 --frames 1 --subtapes 64000 > gpurun_out/r2a/racecheck_prospero.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/run_one.py --model bear --dim 3 --size 128 --frames 1 --subtapes 64000 > gpurun_out/r2a/memcheck_bear.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a/bench_base.json 2> gpurun_out/r2a/bench_base.err
tail -3 gpurun_out/r2a/*.log
