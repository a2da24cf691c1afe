# Round-2 evidence in one go (one GPU): tests, both bench arms + the size sweep, all six models at the reference's
# table sizes, the frame-jitter probe, ncu captures of every kernel of the path, the launch list of a bench run,
# sanitizer logs.  Outputs under gpurun_out/r2ev (copied to profiles/).  ncu reports are exported to CSV on the
# box (raw page for all, source page for the two clause-loop kernels) and not brought back.
set -x
O=gpurun_out/r2ev; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 20 --warmup 5 > $O/bench_mine.json 2> $O/bench_mine.err
python bench.py --impl reference --workload sweep --steps 10 --warmup 3 --no-cpu > $O/sweep_ref.json 2> $O/sweep_ref.err
python bench.py --workload sweep --steps 10 --warmup 3 --no-cpu > $O/sweep_mine.json 2> $O/sweep_mine.err
python tools/sweep.py > $O/sweep_all.log 2>&1; cp gpurun_out/sweep.md $O/sweep_all_models.md; cp gpurun_out/sweep.json $O/sweep_all_models.json
JITTER_FRAMES=320 python tools/frame_jitter.py prospero:2:1024 hello_world:3:512 prospero:2:256 2>&1 | grep "quartiles\|fastest" > $O/frame_jitter.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_voxels -s 2 -c 1 -o $O/prof_voxels python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_voxels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_tiles -s 5 -c 1 -o $O/prof_tiles_L2 python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_tiles.log 2>&1
ncu --set full --clock-control none -k regex:k_normals -s 2 -c 1 -o $O/prof_normals python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_normals.log 2>&1
ncu --set full --clock-control none -k regex:k_eval_root -s 2 -c 1 -o $O/prof_root python tools/run_one.py --model prospero --dim 2 --size 4096 --frames 4 > $O/ncu_root.log 2>&1
ncu --set full --clock-control none -k regex:k_eval_sub -s 2 -c 1 -o $O/prof_sub python tools/run_one.py --model prospero --dim 2 --size 256 --frames 4 > $O/ncu_sub.log 2>&1
ncu --set full --clock-control none -k regex:k_eval_voxels -s 2 -c 1 -o $O/prof_voxels_remap python tools/run_one.py --model involute_gear_3d --dim 3 --size 1024 --frames 4 > $O/ncu_voxels_remap.log 2>&1
for n in voxels tiles_L2 normals root sub voxels_remap; do ncu -i $O/prof_$n.ncu-rep --page raw --csv > $O/raw_$n.csv 2>/dev/null; done
for n in voxels tiles_L2; do ncu -i $O/prof_$n.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/src_$n.csv.gz; done
rm -f $O/*.ncu-rep
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/run_one.py --model bear --dim 3 --size 128 --frames 1 --subtapes 64000 > $O/racecheck_bear_ptxloop.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/run_one.py --model prospero --dim 2 --size 256 --frames 1 --subtapes 64000 > $O/racecheck_prospero_remap_sub.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/run_one.py --model prospero --dim 2 --size 256 --frames 2 --subtapes 64000 > $O/memcheck_prospero_sub.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/run_one.py --model hello_world --dim 3 --size 128 --frames 2 --subtapes 64000 > $O/memcheck_hello_world.log 2>&1
for f in $O/racecheck_bear_ptxloop.log $O/racecheck_prospero_remap_sub.log $O/memcheck_prospero_sub.log $O/memcheck_hello_world.log; do tail -n 2 $f; done
python tools/kernel_times.py bear:3:1024 prospero:2:4096 prospero:2:256 bear:3:256 architecture:3:2048 involute_gear_3d:3:2048 2>&1 | cut -c1-500
ls -la $O
