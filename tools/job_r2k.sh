set -x
O=gpurun_out/r2k; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python tools/kernel_times.py 2>&1 | cut -c1-400 | tee $O/kt.log
# fixtures for the sweep sizes that have none yet
python tools/gpu_check.py --no-cpu --iters 3 --out $O --cases prospero_2d_512,prospero_2d_3072,bear_3d_512 > $O/check.log 2>&1; tail -3 $O/check.log | cut -c1-300
