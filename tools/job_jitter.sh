set -x
JITTER_FRAMES=320 python tools/frame_jitter.py prospero:2:1024 hello_world:3:512 2>&1 | grep -v "wall_ms\|first 96\|n_active" | cut -c1-600
