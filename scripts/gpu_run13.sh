python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --config 5 --cells 40 40 40 --steps 3 --warmup 3 --no-cpu-baseline --md-steps 0 > gpurun_out/bench_c5_r1b.json 2> gpurun_out/bench_c5_r1b.err
python bench.py --config 3 --variant small --steps 20 --warmup 3 --no-cpu-baseline --md-steps 0 > gpurun_out/bench_c3_r1b.json 2> gpurun_out/bench_c3_r1b.err
python bench.py --config 1 --steps 20 --warmup 3 --no-cpu-baseline --md-steps 20 > gpurun_out/bench_c1_r1b.json 2> gpurun_out/bench_c1_r1b.err
python bench.py --config 4 --grade-every 10 --steps 20 --warmup 3 --no-cpu-baseline --md-steps 0 > gpurun_out/bench_c4_r1b.json 2> gpurun_out/bench_c4_r1b.err
python - <<PY
import json
for c in (5,3,1,4):
    try:
        d=json.load(open(f"gpurun_out/bench_c{c}_r1b.json"))
        print("config", c, round(d["value"],2), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}, d.get("md",{}).get("total_energy_drift_eV_per_atom"))
    except Exception as e: print(c, e)
PY
