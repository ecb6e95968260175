#!/bin/bash
# A/B of the dense Cholesky variants at config 5 (radial, n = 12 003) on one box: scripts/ab_chol.sh "chol_big=0" "chol_big=1 chol_chain=0" ...
i=0
for v in "$@"; do
  flags=""; for kv in $v; do flags="$flags --tune $kv"; done
  python bench.py --workload ba_20k_2k --num-intrinsics 3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra $flags \
     > gpurun_out/chol_$i.json 2> gpurun_out/chol_$i.err
  python - "$v" gpurun_out/chol_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k=d['kernels']
    print("%-28s ms/step %.3f cost %.9f  " % (sys.argv[1], d['ms_per_step'], d['config']['final_cost']), {n: round(v['us_per_launch'],1) for n, v in k.items() if v['us_per_launch'] > 50})
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  i=$((i+1))
done
