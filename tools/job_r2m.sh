set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "fixture or live or cpu_restatement" 2>&1 | tail -6
python tools/kernel_times.py bear:3:1024 bear:3:512 hello_world:3:1024 hello_world:2:4096 2>&1 | cut -c1-400
MPRB_FLOAT_TMEM=0 python tools/kernel_times.py bear:3:1024 hello_world:3:1024 2>&1 | cut -c1-300
