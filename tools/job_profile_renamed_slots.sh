set -x
O=gpurun_out/r2q; mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:k_eval_voxels -s 2 -c 1 -o $O/prof_voxels_gears python tools/run_one.py --model involute_gear_3d --dim 3 --size 1024 --frames 4 > $O/ncu_voxels_gears.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval_tiles -s 5 -c 1 -o $O/prof_tiles_gears python tools/run_one.py --model involute_gear_3d --dim 3 --size 1024 --frames 4 > $O/ncu_tiles_gears.log 2>&1
for n in voxels_gears tiles_gears; do
ncu -i $O/prof_$n.ncu-rep --page source --csv --print-source sass > $O/src_$n.csv 2>/dev/null
ncu -i $O/prof_$n.ncu-rep --page raw --csv > $O/raw_$n.csv 2>/dev/null
rm -f $O/prof_$n.ncu-rep
done
gzip -f $O/src_*.csv
ls -la $O
