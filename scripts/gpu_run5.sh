for dbg in 0 32 64 96; do
  MTP_B200_PROG_DEBUG=$dbg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/tmp_dbg.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_dbg.json"))
print("debug $dbg", round(d["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
