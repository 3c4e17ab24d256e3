for c in 32768 65536 98304 131072 262144; do for l in 2 3; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --md-steps 0 --lanes $l --chunksize $c > gpurun_out/tmp_s.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_s.json"))
print("chunk $c lanes $l", round(d["value"],1), round(d["e2e"]["value"],1), round(d["e2e"]["list_resent_every_step"]["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done; done
