"""Development check: numpy restatement of cv2.resize (uint8, INTER_LINEAR, incl. the exact-2x box-average case) against cv2
itself.  The maintained copy is oracle/crop_ref.py resize_linear; the CUDA kernel is csrc/crop_ops.cu k_resize_linear_u8."""
import numpy as np, cv2
def resize_ref(img, dw, dh):
    sh, sw = img.shape[:2]
    cn = img.shape[2]
    inv_x = dw/ sw; inv_y = dh / sh
    sx_scale = 1.0/inv_x; sy_scale = 1.0/inv_y
    def coeffs(dn, sn, scale, clampf=True):
        d = np.arange(dn)
        f = ((d+0.5)*scale - 0.5).astype(np.float32)   # fx = (float)((dx+0.5)*scale_x - 0.5)
        s = np.floor(f).astype(np.int64)
        f = f - s.astype(np.float32)
        lo = s < 0
        hi = s >= sn-1
        if clampf:
            f = np.where(lo, np.float32(0), f); s = np.where(lo, 0, s)
            f = np.where(hi, np.float32(0), f); s = np.where(hi, sn-1, s)
        a0 = np.rint((np.float32(1)-f)*np.float32(2048)).astype(np.int64)
        a1 = np.rint(f*np.float32(2048)).astype(np.int64)
        return s, a0, a1, hi
    sx, ax0, ax1, xhi = coeffs(dw, sw, sx_scale)
    sy, ay0, ay1, yhi = coeffs(dh, sh, sy_scale, False)
    I = img.astype(np.int64)
    sx1 = np.minimum(sx+1, sw-1)
    # horizontal
    Hrow = I[:, sx]*ax0[None,:,None] + I[:, sx1]*ax1[None,:,None]    # [sh, dw, cn]
    sy0 = np.clip(sy, 0, sh-1); sy1 = np.clip(sy+1, 0, sh-1)
    S0 = Hrow[sy0]; S1 = Hrow[sy1]
    out = (((ay0[:,None,None]*(S0>>4))>>16) + ((ay1[:,None,None]*(S1>>4))>>16) + 2) >> 2
    return np.clip(out,0,255).astype(np.uint8)
rng=np.random.default_rng(0); bad=0; tot=0
for t in range(200):
    h,w=int(rng.integers(3,120)),int(rng.integers(3,700))
    img=rng.integers(0,256,(h,w,3),dtype=np.uint8)
    dh=32; dw=max(1,int(32*w/h)); dw=min(dw,804)
    a=cv2.resize(img,(dw,dh)); b=resize_ref(img,dw,dh)
    d=(a!=b); tot+=d.size; bad+=int(d.sum())
    if d.any() and t<30: print(t,h,w,dw,int(d.sum()), int(np.abs(a.astype(int)-b).max()))
print("mismatch",bad,"of",tot)
def resize_full(img,dw,dh):
    sh,sw=img.shape[:2]
    if sw==2*dw and sh==2*dh:
        I=img.astype(np.int64)
        return ((I[0::2,0::2]+I[0::2,1::2]+I[1::2,0::2]+I[1::2,1::2]+2)>>2).astype(np.uint8)
    return resize_ref(img,dw,dh)
rng=np.random.default_rng(1); bad=0; tot=0
for t in range(300):
    h,w=int(rng.integers(2,200)),int(rng.integers(2,1500))
    if t%10==0: h=64; w=2*int(rng.integers(10,400))
    img=rng.integers(0,256,(h,w,3),dtype=np.uint8)
    ratio=w/float(h)
    dw=804 if ratio>804/32 else int(32*ratio)
    if dw<1: continue
    a=cv2.resize(img,(dw,32)); b=resize_full(img,dw,32)
    d=(a!=b); tot+=d.size; bad+=int(d.sum())
    if d.any(): print("bad",t,h,w,dw,int(d.sum()))
print("keepratio mismatch",bad,"of",tot)
# pp-ocr rec: height 48, arbitrary widths
for t in range(200):
    h,w=int(rng.integers(2,150)),int(rng.integers(2,900))
    img=rng.integers(0,256,(h,w,3),dtype=np.uint8)
    dw=int(rng.integers(16,1281))
    a=cv2.resize(img,(dw,48)); b=resize_full(img,dw,48)
    d=(a!=b); tot+=d.size; bad+=int(d.sum())
    if d.any(): print("bad48",t,h,w,dw,int(d.sum()))
print("total mismatch",bad,"of",tot)
