set -x
O=gpurun_out/r2n; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_one.py --model hello_world --dim 3 --size 128 --frames 2 --subtapes 64000 2>&1 | grep -v "Saved host\|Host Frame\|^=========         " | tail -5
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/run_one.py --model bear --dim 3 --size 128 --frames 1 --subtapes 64000 2>&1 | grep -v "Saved host\|Host Frame\|^=========         " | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2n/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["config"]["ms_per_frame"], d["frames_verified"], d["kernel_ms_per_step"], d["config"]["effects_ms"], d["roofline"]["frac"])
PY
