set -x
O=gpurun_out/r2h; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.err
python bench.py --steps 10 --warmup 3 > $O/bench_mine.json 2> $O/bench_mine.err; tail -c 600 $O/bench_mine.err
python - <<'PY'
import json
for f in ("bench_ref","bench_mine"):
    try:
        d=json.load(open(f"gpurun_out/r2h/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["config"]["ms_per_frame"], d["config"].get("wall_ms_per_frame"), d.get("frames_verified"), d.get("frames_verified_detail"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_baseline",{}).get("sample","")[:300])
    except Exception as e:
        print(f, "ERR", e)
PY
