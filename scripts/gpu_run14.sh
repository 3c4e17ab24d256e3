python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for l in 1 2; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --md-steps 0 --lanes $l > gpurun_out/tmp_g.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/tmp_g.json"))
print("lanes $l", round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
