#!/bin/bash
# A/B of kernel variants on one box: scripts/ab_bench.sh <tag> "<tune flags 1>" "<tune flags 2>" ...
tag=$1; shift
i=0
for flags in "$@"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $flags > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$flags" gpurun_out/${tag}_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k=d['kernels']
    print("%-44s ms/step %.4f  E %.1f F %.1f schur %.1f pcg %.1f backsub %.1f cand %.1f cost %.6f" % (sys.argv[1], d['ms_per_step'], k['accum_E']['us_per_launch'], k['accum_F']['us_per_launch'], k['schur_eliminate']['us_per_launch'], k.get('pcg_solve',{'us_per_launch':0})['us_per_launch'], k['backsub']['us_per_launch'], k['candidate']['us_per_launch'], d['config']['final_cost']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  i=$((i+1))
done
