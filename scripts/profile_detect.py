"""Two passes of the marker detector over the bench's 64-frame batch (frames resident in HBM), for ncu:
  ncu --set full --clock-control none --import-source on -k regex:ard -c 12 -o gpurun_out/detect python scripts/profile_detect.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ar_slam_b200 import capi, synth  # noqa: E402

n, h, w = int(os.environ.get("FRAMES", "64")), 768, 1020
bits = synth.dict_4x4_50_bits()
distinct = [synth.render_marker_scene(h, w, bits, 8, 0xA55A0600 + k, noise=4.0)[0] for k in range(8)]
dev = torch.stack([torch.from_numpy(distinct[i % 8]) for i in range(n)]).cuda()
torch.cuda.synchronize()
det = capi.Detector(n, w, h)
p = capi.default_detect_params(min_corner_distance_rate=0.1)
for _ in range(2):
    res = det.detect(None, p, device_ptr=dev.data_ptr(), shape=dev.shape)
    print(det.times(), sum(len(i) for i, _ in res))
