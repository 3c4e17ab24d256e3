set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_p3a.log
cat gpurun_out/pytest_p3a.log
for cfg in "2 1" "2 2"; do
  set -- $cfg
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes $2 > gpurun_out/bench_p3a_l$2.json 2> gpurun_out/bench_p3a_l$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_p3a_l$2.json"))
print("lanes $2", d["value"], d["e2e"]["value"], {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
MTP_B200_NO_PROG_V3=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 2 > gpurun_out/bench_p3a_old.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_p3a_old.json'))
print('old', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
MTP_B200_PROG_DSMEM=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 2 > gpurun_out/bench_p3a_l2ds.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_p3a_l2ds.json'))
print('l2 dsmem', d['value'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
