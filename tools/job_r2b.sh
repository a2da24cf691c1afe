set -x
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q > $O/pytest_parity.log 2>&1; tail -5 $O/pytest_parity.log
for G in 1 2 4; do
  MPRB_FLOAT_GROUP=$G python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_g$G.json 2> $O/bench_g$G.err
  python - <<PY
import json
d=json.load(open("$O/bench_g$G.json"))
print("G=$G", d["value"], d["config"]["ms_per_frame"], d["kernel_ms_per_step"])
PY
done
for G in 1 2 4; do MPRB_FLOAT_GROUP=$G python tools/kernel_times.py > $O/kt_g$G.json 2>&1; done
