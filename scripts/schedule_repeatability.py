"""How repeatable is the reference's incremental schedule (solve(), ar_slam_util.cpp:744-866) from f0 = 3000 on a synthetic
map?  Three maps x two data flows (parameters resident on the GPU / moved around every optimize()) x three runs each.
The first solves (one capture, its tags and the focal length all free, 50-iteration cap) are so under-determined that the
last bits decide where they wander; the FP64 reductions of the Schur scatter arrive in a different order in every run, so
the same binary on the same input ends in different minima (profiles/r2_schedule_repeatability.txt)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ar_slam_b200 import synth
CLI = os.path.join(ROOT, "ar_slam_b200", "lib", "ar_slam_cli")
tmp = tempfile.mkdtemp()
for seed in (400, 401, 402):
    m = synth.make_map(200, 50, 8, seed=0xA55A0000 + seed)
    det = os.path.join(tmp, "d%d.yaml" % seed)
    synth.write_detections_yaml(m, det)
    for name, flags in (("device", []), ("host", ["--host-params"])):
        for rep in range(3):
            r = subprocess.run([CLI, "--quiet", "--output", "o.yaml"] + flags + [det], cwd=tmp, capture_output=True, text=True)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("schedule:")][0]
            cam = [ln for ln in r.stdout.splitlines() if "f=" in ln][0]
            print(seed, name, rep, line.split("final cost")[1].strip(), cam.strip(), line.split(",")[1].strip(), flush=True)
