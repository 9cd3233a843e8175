import sys, numpy as np, cv2
sys.path.insert(0,'/root/repo')
from tests import cases
w,h=int(sys.argv[1]),int(sys.argv[2])
prev,curr = cases.make_gray_pair(dict(w=w,h=h,seed=w+h,amount=1.0))
d = cv2.DISOpticalFlow.create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
d.setFinestScale(2); d.setPatchSize(8); d.setPatchStride(4); d.setUseSpatialPropagation(True)
try:
    ref=d.calc(prev,curr,None); print(w,h,'ok',d.getFinestScale(), float(np.abs(ref).max()), np.isfinite(ref).all())
    ref=d.calc(prev,curr,None); print(w,h,'ok2',d.getFinestScale(), float(np.abs(ref).max()))
except cv2.error as e:
    print(w,h,'cv2.error',str(e)[-100:].strip())
