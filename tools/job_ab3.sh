set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
JITTER_FRAMES=640 python tools/frame_jitter.py prospero:2:1024 hello_world:3:512 prospero:2:256 2>&1 | grep -v "wall_ms" | cut -c1-560
python tools/sweep.py 2>&1 | tail -0
cat gpurun_out/sweep.md
