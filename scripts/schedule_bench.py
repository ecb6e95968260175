"""Wall time of the reference's real workload -- the incremental map build, one optimize() per added capture
(solve(), ar_slam_util.cpp:744-866) -- through the drop-in CLI on the GPU, against the oracle walking the same
schedule on the host cores.  Writes one JSON document (profiles/r2_schedule_bench.json is a copy of it).

  python scripts/schedule_bench.py > gpurun_out/r2_schedule_bench.json
"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ar_slam_b200 import synth  # noqa: E402
from oracle import schedule  # noqa: E402

CLI = os.path.join(ROOT, "ar_slam_b200", "lib", "ar_slam_cli")
GOLD = os.path.join(ROOT, "tests", "golden")


def run_cli(det, flags, cwd):
    t = time.perf_counter()
    r = subprocess.run([CLI, "--quiet", "--output", "out.yaml"] + flags + [det], cwd=cwd, capture_output=True, text=True)
    wall = time.perf_counter() - t
    if r.returncode != 0:
        raise RuntimeError(r.stdout[-1000:] + r.stderr[-1000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("schedule:")][0].split()
    cam = [ln for ln in r.stdout.splitlines() if "f=" in ln][0]
    return {"schedule_s": float(line[15]) * 1e-3, "process_wall_s": wall, "solves": int(line[1]), "lm_iterations": int(line[3]),
            "ms_in_arslam_solve": float(line[6]), "ms_on_device": float(line[10].lstrip("(")), "final_cost": float(line[-1]),
            "focal": float(cam.split("f=")[1].split()[0])}


def oracle_schedule(det):
    m = schedule.MapData()
    m.load_yaml(det)
    t = time.perf_counter()
    schedule.Scheduler(m).solve()
    return {"schedule_s": time.perf_counter() - t, "solves": len(m.solve_log), "lm_iterations": int(sum(s["iterations"] for s in m.solve_log)),
            "final_cost": m.solve_log[-1]["final_cost"], "focal": float(m.cam[0]), "threads": 1,
            "what": "oracle/schedule.py: the restated schedule + restated ceres::Solve (Jets, dense Schur), one thread like the reference"}


def main():
    out = {"what": __doc__.strip().splitlines()[0]}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- BASELINE config 1: the demo map (3 captures, 6 tags)
        demo = os.path.join(GOLD, "demo_map_detections.yaml")
        run_cli(demo, [], tmp)   # warm the driver / context once
        out["demo_map"] = {"gpu_device_resident": run_cli(demo, [], tmp), "gpu_host_params": run_cli(demo, ["--host-params"], tmp),
                           "cpu_oracle": oracle_schedule(demo)}
        for name, (nc, nt, tpc, with_cpu) in (("synthetic_200x50", (200, 50, 8, True)), ("synthetic_1000x200", (1000, 200, 8, False))):
            m = synth.make_map(nc, nt, tpc, seed=0xA55A0000 + 200 + nc)
            det = os.path.join(tmp, name + ".yaml")
            synth.write_detections_yaml(m, det)
            rec = {"captures": nc, "tags": nt, "blocks": int(len(m.cap_idx)),
                   "gpu_device_resident": run_cli(det, [], tmp),
                   "gpu_host_params": run_cli(det, ["--host-params"], tmp),
                   "gpu_8_captures_per_solve": run_cli(det, ["--captures-per-solve", "8"], tmp)}
            if with_cpu:
                rec["cpu_oracle"] = oracle_schedule(det)
            out[name] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
