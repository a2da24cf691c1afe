set -x
python tools/kernel_times.py prospero:2:256 prospero:2:512 prospero:2:1024 hello_world:2:1024 bear:3:256 bear:3:1024 prospero:2:4096 2>&1 | cut -c1-1200
MPRB_KT_SHARD=8:3 python tools/kernel_times.py bear:3:1024 prospero:2:4096 2>&1 | cut -c1-1200
python tools/run_one.py --impl ref --model hello_world --dim 2 --size 1024 --frames 3
