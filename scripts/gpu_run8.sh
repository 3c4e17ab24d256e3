python -m pytest tests/test_gpu_neigh.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nb1.json 2> gpurun_out/bench_nb1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_nb1.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), d["e2e"]["list_resent_every_10th_step"]["value"], d["e2e"]["device_built_list"])
PY
tail -3 gpurun_out/bench_nb1.err
