import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ar_slam_b200 import synth
from oracle import schedule
CLI = os.path.join(ROOT, "ar_slam_b200", "lib", "ar_slam_cli")
m = synth.make_map(200, 50, 8, seed=0xA55A0000 + 400)
tmp = tempfile.mkdtemp()
det = os.path.join(tmp, "d.yaml")
synth.write_detections_yaml(m, det)
logs = {}
for name, flags in (("device", []), ("host", ["--host-params"])):
    r = subprocess.run([CLI, "--quiet", "--solve-log", "--output", "o.yaml"] + flags + [det], cwd=tmp, capture_output=True, text=True)
    logs[name] = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("solve ")]
ref = schedule.MapData(); ref.load_yaml(det); schedule.Scheduler(ref).solve()
for i in range(len(ref.solve_log)):
    o = ref.solve_log[i]
    d, h = logs["device"][i], logs["host"][i]
    print(i, "oracle it %d %.6g -> %.6g | device it %s %s -> %s | host it %s %s -> %s" % (o["iterations"], o["initial_cost"], o["final_cost"], d[3], d[5], d[7], h[3], h[5], h[7]))
