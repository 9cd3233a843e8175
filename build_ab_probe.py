import ctypes as C, numpy as np, cv2, sys
sys.path.insert(0,'/root/repo')
from oracle import dis_ref
from tests import cases
L = dis_ref.lib()
def cvdis():
    d = cv2.DISOpticalFlow.create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
    d.setFinestScale(2); d.setPatchSize(8); d.setPatchStride(4); d.setUseSpatialPropagation(True)
    return d
def scales(i0,i1,f,c):
    h,w=i0.shape; flow=np.zeros((h,w,2),np.float32)
    rc=L.disref_calc_scales(dis_ref._p(i0),dis_ref._p(i1),C.c_int(h),C.c_int(w),C.byref(dis_ref.default_params()),C.c_int(f),C.c_int(c),dis_ref._p(flow))
    return rc, flow
def sel(h,w,finest):
    fi=C.c_int(); co=C.c_int()
    rc=L.disref_select_scales(C.c_int(h),C.c_int(w),C.byref(dis_ref.default_params(finest_scale=finest)),C.byref(fi),C.byref(co))
    return rc,fi.value,co.value
bad=0
for (w,h) in [(73,45),(121,73),(64,48),(90,50),(91,52),(100,32),(48,32),(40,24),(32,32),(24,16),(16,12),(45,73),(30,100),(12,12),(20,9),(128,72),(160,90),(64,20),(89,33),(90,90),(33,91),(57,57),(70,40),(80,45), (85, 48), (64,36),(96,54),(72,72)]:
    for amount in (1.0, 3.0):
        prev,curr = cases.make_gray_pair(dict(w=w,h=h,seed=w+h,amount=amount))
        d=cvdis()
        ref=d.calc(prev,curr,None); ref2=d.calc(curr,prev,None); ref3=d.calc(prev,curr,None)
        rc,f1,c1=sel(h,w,2); _,fl1=scales(prev,curr,f1,c1)
        rc2,f2,c2=sel(h,w,f1); _,fl2=scales(curr,prev,f2,c2); _,fl3=scales(prev,curr,f2,c2)
        ok=np.array_equal(ref,fl1) and np.array_equal(ref2,fl2) and np.array_equal(ref3,fl3)
        bad+= not ok
        print(w,h,amount,'first',(f1,c1),'later',(f2,c2),'OK' if ok else 'MISMATCH', float(np.abs(ref).max()))
print('bad',bad)
