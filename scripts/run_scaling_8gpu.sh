set -x
timeout 300 python -m pytest tests/test_gpu_multi.py -q -k "overlap or (direct and 8) or (staged and 4)" > gpurun_out/r2_multi8.log 2>&1; tail -3 gpurun_out/r2_multi8.log
run() { # n config extra...
  n=$1; c=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$c bench.py --gpus $n --config $c --steps 20 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/err_$n_$c.log
}
run 8 2 > gpurun_out/r2_c2_n8.json; run 4 2 > gpurun_out/r2_c2_n4.json
run 8 2 --overlap > gpurun_out/r2_c2_n8_overlap.json
run 8 2 --overlap --chunksize 49152 > gpurun_out/r2_c2_n8_overlap_c48.json
run 8 5 --steps 5 --warmup 3 > gpurun_out/r2_c5_n8.json; run 4 5 --steps 5 --warmup 3 > gpurun_out/r2_c5_n4.json
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_c*_n[48]*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],2), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["list_resident"]["value"],2))
    except Exception as e: print(f, "ERR", e)
PY
