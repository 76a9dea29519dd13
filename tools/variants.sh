#!/bin/bash
# usage: tools/variants.sh <tag> "<ENV1=.. ENV2=..>" ...   — runs the 5 %-scale bench once per environment string and prints the stage times
tag=$1; shift
i=0
for envs in "$@"; do
  out=gpurun_out/${tag}_$i.json
  env $envs python bench.py --scale ${SCALE:-0.05} --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > $out 2> gpurun_out/${tag}_$i.err
  grep "rcb tile phases\|rcb merge phases" gpurun_out/${tag}_$i.err
  python - "$out" "$envs" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2],"| ms/step %.3f |"%d["ms_per_step"]," ".join("%s=%.3f"%(k.split("(")[0],v) for k,v in d["stage_ms"].items()))
except Exception as e:
    print(sys.argv[2],"FAILED",e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
  i=$((i+1))
done
