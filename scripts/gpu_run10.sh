for na in 32 16; do for l in 1 2 3; do
  MTP_B200_P3_NA=$na python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes $l > gpurun_out/tmp_na.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_na.json"))
print("na $na lanes $l", round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done; done
