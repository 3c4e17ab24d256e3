# phase split of the v3 program kernel (timing experiments) + one ncu capture
for dbg in 0 1 2 3 7 31 24; do
  MTP_B200_PROG_DEBUG=$dbg python bench.py --steps 5 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/tmp_dbg.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_dbg.json"))
print("debug $dbg", round(d["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mtp_program_v3 -s 8 -c 1 -o gpurun_out/prof_p3a -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/ncu_p3a.log 2>&1
ls -la gpurun_out/prof_p3a.ncu-rep
