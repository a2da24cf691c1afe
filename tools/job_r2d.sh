set -x
O=gpurun_out/r2d; mkdir -p $O
./build/ubench_dispatch > $O/ubench.log 2>&1; cat $O/ubench.log
python -m pytest tests/test_gpu_parity.py -q -k "effects" 2>&1 | tail -30
