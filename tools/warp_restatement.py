"""Development check: numpy restatement of cv2.warpPerspective (uint8, INTER_LINEAR, BORDER_CONSTANT 0) against cv2 itself
on 300 random quads.  The maintained copy of the restatement is oracle/crop_ref.py; the CUDA kernel is csrc/crop_ops.cu."""
import numpy as np, cv2, math
print(cv2.__version__)
def warp_ref(img, T, w, h):
    """numpy restatement of cv2.warpPerspective(img, T, (w,h)) INTER_LINEAR BORDER_CONSTANT 0, uint8 HWC."""
    M = cv2.invert(T)[1].astype(np.float64).ravel()
    sh, sw = img.shape[:2]
    BLOCK=32
    bh0=min(BLOCK//2,h); bw0=min(BLOCK*BLOCK//bh0,w); bh0=min(BLOCK*BLOCK//bw0,h)
    ys=np.arange(h,dtype=np.float64)[:,None]
    xs=np.arange(w)
    xb=(xs//bw0*bw0).astype(np.float64)[None,:]
    x1=(xs%bw0).astype(np.float64)[None,:]
    X0=M[0]*xb+M[1]*ys+M[2]
    Y0=M[3]*xb+M[4]*ys+M[5]
    W0=M[6]*xb+M[7]*ys+M[8]
    W=W0+M[6]*x1
    with np.errstate(divide='ignore',invalid='ignore'):
        Wi=np.where(W!=0, 32.0/W, 0.0)
    fX=np.maximum(-2147483648.0,np.minimum(2147483647.0,(X0+M[0]*x1)*Wi))
    fY=np.maximum(-2147483648.0,np.minimum(2147483647.0,(Y0+M[3]*x1)*Wi))
    X=np.rint(fX).astype(np.int64); Y=np.rint(fY).astype(np.int64)
    sx=np.clip(X>>5,-32768,32767); sy=np.clip(Y>>5,-32768,32767)
    ax=(X&31); ay=(Y&31)
    out=np.zeros((h,w,img.shape[2]),np.uint8)
    w00=(32-ax)*(32-ay)*32; w01=ax*(32-ay)*32; w10=(32-ax)*ay*32; w11=ax*ay*32
    def fetch(yy,xx):
        ok=(yy>=0)&(yy<sh)&(xx>=0)&(xx<sw)
        v=img[np.clip(yy,0,sh-1),np.clip(xx,0,sw-1)].astype(np.int64)
        return np.where(ok[...,None],v,0)
    acc=fetch(sy,sx)*w00[...,None]+fetch(sy,sx+1)*w01[...,None]+fetch(sy+1,sx)*w10[...,None]+fetch(sy+1,sx+1)*w11[...,None]
    return ((acc+(1<<14))>>15).astype(np.uint8)
rng=np.random.default_rng(0)
bad=0; tot=0
for trial in range(300):
    H,W=rng.integers(100,700),rng.integers(100,900)
    img=rng.integers(0,256,(H,W,3),dtype=np.uint8)
    cx,cy=rng.uniform(50,W-50),rng.uniform(50,H-50)
    bw,bh=rng.uniform(10,400),rng.uniform(6,80)
    ang=rng.uniform(-0.6,0.6)
    c,s=math.cos(ang),math.sin(ang)
    pts=np.array([[-bw/2,-bh/2],[bw/2,-bh/2],[-bw/2,bh/2],[bw/2,bh/2]])@np.array([[c,s],[-s,c]])+[cx,cy]
    pts=pts+rng.uniform(-3,3,pts.shape)
    corners=pts.astype(np.float32)
    iw=math.dist(((pts[0]+pts[2])/2),((pts[1]+pts[3])/2)); ih=math.dist((pts[0]+pts[1])/2,(pts[2]+pts[3])/2)
    ct=np.array([[0,0],[iw-1,0],[0,ih-1],[iw-1,ih-1]],np.float32)
    T=cv2.getPerspectiveTransform(corners,ct)
    w,h=int(iw),int(ih)
    if w<1 or h<1: continue
    a=cv2.warpPerspective(img,T,(w,h))
    b=warp_ref(img,T,w,h)
    d=(a!=b).any(-1)
    tot+=d.size; bad+=int(d.sum())
    if d.any() and bad<2000 and trial<40: print(trial,w,h,int(d.sum()), np.argwhere(d)[:3].tolist())
print("mismatch pixels",bad,"of",tot)
