import numpy as np, cv2, sys, time
sys.path.insert(0,'/root/repo')
from oracle import resample_np as R
rng=np.random.default_rng(5)
cv2.setNumThreads(1)
bad=0; worst={'bilinear':0,'bicubic':0}; maskbad=0
for k in range(120):
    w=int(rng.integers(8,400)); h=int(rng.integers(8,300))
    ow=int(w+rng.integers(-5,40)); oh=int(h+rng.integers(-5,40))
    src=rng.random((h,w,3),dtype=np.float32)
    th=rng.normal(0,0.05); sc=1+rng.normal(0,0.05)
    M=np.array([[sc*np.cos(th),-sc*np.sin(th),rng.normal(0,8)],[sc*np.sin(th),sc*np.cos(th),rng.normal(0,8)],[0,0,1]],np.float64)
    if k%3==0: M[2,:2]=rng.normal(0,2e-4,2)
    if k%10==0: M=np.array([[1,0,float(rng.integers(-4,5))+ (0.5 if k%20==0 else 0)],[0,1,float(rng.integers(-4,5))],[0,0,1]],np.float64)
    M=M.astype(np.float32)
    border=tuple(float(x) for x in (rng.integers(0,256,3)/255.0).astype(np.float32))
    for interp,flag in (('bilinear',cv2.INTER_LINEAR),('bicubic',cv2.INTER_CUBIC)):
        ref=cv2.warpPerspective(src,M,(ow,oh),flags=flag,borderMode=cv2.BORDER_CONSTANT,borderValue=border)
        mine=R.warp_np(src,M,(ow,oh),interp,border)
        e=float(np.abs(ref-mine).max()); worst[interp]=max(worst[interp],e)
        if interp=='bilinear' and e!=0: bad+=1; print('bilinear mismatch',k,w,h,ow,oh,e,int((ref!=mine).sum()))
    ones=np.ones((h,w),np.float32)
    cov=cv2.warpPerspective(ones,M,(ow,oh),flags=cv2.INTER_NEAREST,borderMode=cv2.BORDER_CONSTANT,borderValue=0)>0.5
    mine=R.coverage_np(M,(w,h),(ow,oh))
    d=int((cov!=mine).sum())
    if d: maskbad+=1; print('mask mismatch',k,w,h,ow,oh,d)
print('bilinear bad',bad,'worst',worst,'mask bad',maskbad)
