set -x
O=gpurun_out/r2prof; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_voxels -s 2 -c 1 -o $O/prof_voxels python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_voxels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_tiles -s 5 -c 1 -o $O/prof_tiles_L2 python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_tiles.log 2>&1
ncu --set full --clock-control none -k regex:k_normals -s 2 -c 1 -o $O/prof_normals python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_normals.log 2>&1
ncu --set full --clock-control none -k regex:k_eval_root -s 2 -c 1 -o $O/prof_root python tools/run_one.py --model prospero --dim 2 --size 4096 --frames 4 > $O/ncu_root.log 2>&1
ls -la $O
