"""GPU detector against the CPU restatement on frames whose width is not a multiple of 4 or 8 (ad-hoc check)."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from ar_slam_b200 import capi, synth
from oracle import aruco_detect as A
bits=synth.dict_4x4_50_bits()
for (h,w) in ((251,333),(97,131)):
    img=synth.render_marker_scene(h,w,bits,3,5,noise=3.0)[0]
    det=capi.Detector(1,w,h)
    ids,c=det.detect(img[None],capi.default_detect_params())[0]
    oc,oi=A.detect_markers(img,bits,1,A.DEFAULTS)
    g=det.read_stage(0).reshape(h,w); m=det.read_stage(1).reshape(h,w)
    grey=A.to_gray(img)
    ok_mask=all(((((m>>k)&1)!=0)==(A.adaptive_threshold(grey,win,7.0)!=0)).all() for k,win in enumerate((3,13,23)))
    print(h,w,'grey',(g==grey).all(),'mask',ok_mask,'ids',list(map(int,ids)),oi,'corners',all((a==b).all() for a,b in zip(c,oc)) and len(c)==len(oc))
