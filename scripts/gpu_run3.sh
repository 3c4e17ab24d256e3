python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for l in 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes $l > gpurun_out/bench_p3b_l$l.json 2>gpurun_out/bench_p3b_l$l.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_p3b_l$l.json"))
print("lanes $l", round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
for dbg in 3 31; do
  MTP_B200_PROG_DEBUG=$dbg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/tmp_dbg.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_dbg.json"))
print("debug $dbg", round(d["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
MTP_B200_PROG_DSMEM=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/tmp_dbg.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/tmp_dbg.json"))
print("l1 nodsmem", round(d["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
