#!/bin/bash
# usage (under gpurun --gpus N): scripts/run_ngpu.sh N   -> gpurun_out/r2_mgpu_check_N.log, r2_bench_ba_100k_5k_Ngpu.json, r2_bench_loc_1m_5k_Ngpu.json
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/mgpu_check.py > gpurun_out/r2_mgpu_check_${N}.log 2>&1
grep -E "OK|MISMATCH" gpurun_out/r2_mgpu_check_${N}.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/r2_bench_ba_100k_5k_${N}gpu.json 2> gpurun_out/r2_bench_ba_100k_5k_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_ba_100k_5k_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --workload loc_1m_5k --steps 5 --warmup 2 > gpurun_out/r2_bench_loc_1m_5k_${N}gpu.json 2> gpurun_out/r2_bench_loc_1m_5k_${N}gpu.err
python - <<PY
import json
for f in ("gpurun_out/r2_bench_ba_100k_5k_${N}gpu.json", "gpurun_out/r2_bench_loc_1m_5k_${N}gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("n_gpu_check"), (d.get("extra") or {}).get("weak", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
