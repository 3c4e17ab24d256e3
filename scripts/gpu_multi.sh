N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r1b_n$N.json 2> gpurun_out/bench_r1b_n$N.err
tail -c 600 gpurun_out/bench_r1b_n$N.json | head -c 600; echo
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r1b_n$N.json"))
print("N=$N", round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"])
PY
[ "$2" = "test" ] && python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
