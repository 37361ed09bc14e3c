import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"])
