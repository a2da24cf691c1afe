set -x
O=gpurun_out/r2p; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:k_eval_voxels -s 2 -c 1 -o $O/prof_voxels python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_voxels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_normals -s 2 -c 1 -o $O/prof_normals python tools/run_one.py --model bear --dim 3 --size 1024 --frames 4 > $O/ncu_normals.log 2>&1
ls -la $O
