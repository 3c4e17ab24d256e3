# One-GPU measurement set of a round: GPU tests, the bench line of every BASELINE config, the reference arm, the ncu
# launch list and one ncu --set full capture of the pipeline kernels.  Usage: gpurun -- bash scripts/run_round_measurements.sh TAG
TAG=${1:-r2}
O=gpurun_out
python -m pytest tests -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; tail -2 $O/${TAG}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
python bench.py --impl reference --steps 8 --warmup 1 > $O/${TAG}_bench_reference.json 2>/dev/null
python bench.py --config 1 --steps 50 --warmup 10 --md-steps 0 > $O/${TAG}_bench_c1.json 2>/dev/null
python bench.py --config 3 --variant small --steps 50 --warmup 10 --md-steps 0 > $O/${TAG}_bench_c3.json 2>/dev/null
python bench.py --config 4 --grade-every 10 --steps 20 --warmup 5 --md-steps 0 > $O/${TAG}_bench_c4.json 2>/dev/null
python bench.py --config 5 --steps 3 --warmup 3 > $O/${TAG}_bench_c5.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --md-steps 0 --lanes 1"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/ncu1.log 2>&1
$CMD > $O/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"mtp_gather_radial|mtp_moments_v2|mtp_program_p4|mtp_forces_v2" -s 8 -c 8 -o $O/${TAG}_kernels $CMD > $O/ncu2.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],3), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],3), round(d["e2e"]["list_resident"]["value"],3), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
