# 8-GPU box: multi-GPU parity tests, then the weak-scaling lines of config 2 and config 5 at N = 1, 2, 4, 8.
# Usage: gpurun --gpus 8 -- bash scripts/run_scaling_8gpu.sh TAG
TAG=${1:-r2}
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -k "overlap or (direct and 8) or (staged and 4)" > $O/${TAG}_pytest_gpu_multi.log 2>&1; tail -3 $O/${TAG}_pytest_gpu_multi.log
run() { # n config extra...
  n=$1; c=$2; shift 2
  if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 --config $c --no-cpu-baseline --md-steps 0 "$@" 2> $O/err_${n}_$c.log
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$c bench.py --gpus $n --config $c --no-cpu-baseline "$@" 2> $O/err_${n}_$c.log; fi
}
for n in 1 2 4 8; do run $n 2 --steps 20 --warmup 5 > $O/${TAG}_scale_c2_n$n.json; done
run 8 2 --steps 20 --warmup 5 --no-overlap > $O/${TAG}_scale_c2_n8_no_overlap.json
for n in 1 8; do run $n 5 --steps 4 --warmup 3 > $O/${TAG}_scale_c5_n$n.json; done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${TAG}_scale_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],2), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["list_resident"]["value"],2))
    except Exception as e: print(f, "ERR", e)
PY
