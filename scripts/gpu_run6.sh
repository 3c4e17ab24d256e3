python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for c in 32768 65536; do for l in 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes $l --chunksize $c > gpurun_out/bench_p3d_c${c}_l$l.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_p3d_c${c}_l$l.json"))
print("chunk $c lanes $l", round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done; done
for dbg in 3 31; do
  MTP_B200_PROG_DEBUG=$dbg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/tmp_dbg.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_dbg.json"))
print("debug $dbg", round(d["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
