# ncu evidence for profiles/: launch list (per-kernel device time) + one --set full capture of each pipeline kernel.
# Each ncu pass follows a plain run of the same command that exited 0.
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 1 --md-steps 0"
$CMD > gpurun_out/bench_${TAG}_plain.json 2> gpurun_out/bench_${TAG}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_${TAG}_launches.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"mtp_(gather_kernel|radial_kernel|moments_v2|program_v3|forces_v2)" -s 60 -c 5 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_${TAG}_full.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep gpurun_out/launches_${TAG}.csv
