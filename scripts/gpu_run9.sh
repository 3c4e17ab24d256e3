# moment-kernel tile experiment: 8 pairs per staged tile (4 CTAs per SM) against 16
for lib in default nt8; do
  if [ $lib = nt8 ]; then cp lammps-mtp-kokkos_b200/libmtp_b200.so /tmp/keep.so; cp exp/libmtp_b200_nt8.so lammps-mtp-kokkos_b200/libmtp_b200.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tmp_nt.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/tmp_nt.json"))
print("$lib", round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
PY
done
