import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from aruco_b200 import MarkerDetector, synth
from oracle import native
from oracle.cv2_oracle import Params
W,H,n=3840,2160,8
frames=np.stack([synth.render_frame(W,H,100,seed=2000+i,sigma=4.0)[0] for i in range(n)])
K,D=synth.camera_for(W,H)
det=MarkerDetector()
t=time.time(); res=det.detect_batch(frames,K,D,0.05); dt=time.time()-t
print("sigma 4: markers/frame", [len(r) for r in res], "counters/frame", {k: v//n for k,v in det.counters().items()}, "ms", round(dt*1e3))
ref=native.detect(frames[0], Params(), K, D, 0.05, debug=False)["markers"]
print("frame0 ids equal oracle:", [m.id for m in res[0]]==[m["id"] for m in ref], len(ref))
