python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_md1.json 2> gpurun_out/bench_md1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_md1.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), d["e2e"]["device_built_list"]["value"], d.get("md"))
PY
tail -3 gpurun_out/bench_md1.err
