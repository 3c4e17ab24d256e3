python bench.py > gpurun_out/bench_r1d_n1.json 2> gpurun_out/bench_r1d_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r1d_n1.json"))
print("N=1", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "every-step", round(d["e2e"]["list_resent_every_step"]["value"],1), "devlist", round(d["e2e"]["device_built_list"]["value"],1), "md", round(d["md"]["value"],1), d["md"]["total_energy_drift_eV_per_atom"])
print(d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline"]["traffic"], {k:(round(v["ms_per_step"],3), round(v.get("frac_of_fp64_peak",0),3)) for k,v in d["roofline"]["kernels"].items()}, round(d["roofline"]["pipeline"]["frac"],3))
print(d["cpu_baseline"]["value"], d["clocks"], d["gpu_launches"], d["config"]["chunksize"])
PY
