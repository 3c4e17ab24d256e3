import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()})
