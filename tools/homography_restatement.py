"""Development check: cv2.getPerspectiveTransform restated (float32 products, LU with partial pivoting in double) against cv2
itself; the inverse (closed-form 3x3) was checked the same way.  CUDA: csrc/crop_ops.cu k_quad_homography."""
import numpy as np, cv2, math
def lu_solve(A,b):
    A=A.copy(); b=b.copy(); m=8
    for i in range(m):
        k=i
        for j in range(i+1,m):
            if abs(A[j,i])>abs(A[k,i]): k=j
        if k!=i:
            A[[i,k],i:]=A[[k,i],i:]; b[[i,k]]=b[[k,i]]
        d=-1/A[i,i]
        for j in range(i+1,m):
            alpha=A[j,i]*d
            for kk in range(i+1,m): A[j,kk]+=alpha*A[i,kk]
            b[j]+=alpha*b[i]
    for i in range(m-1,-1,-1):
        s=b[i]
        for kk in range(i+1,m): s-=A[i,kk]*b[kk]
        b[i]=s/A[i,i]
    return b
def gpt_f32prod(src,dst):
    src=np.asarray(src,np.float32); dst=np.asarray(dst,np.float32)
    A=np.zeros((8,8)); b=np.zeros(8)
    for i in range(4):
        sx,sy,dx,dy=src[i,0],src[i,1],dst[i,0],dst[i,1]   # float32 scalars
        A[i,0]=A[i+4,3]=sx; A[i,1]=A[i+4,4]=sy; A[i,2]=A[i+4,5]=1
        A[i,6]=np.float32(-sx*dx); A[i,7]=np.float32(-sy*dx); A[i+4,6]=np.float32(-sx*dy); A[i+4,7]=np.float32(-sy*dy)
        b[i]=dx; b[i+4]=dy
    return np.append(lu_solve(A,b),1.0).reshape(3,3)
rng=np.random.default_rng(0); ok=0; n=0
for t in range(500):
    cx,cy=rng.uniform(0,960),rng.uniform(0,960); bw,bh,ang=rng.uniform(4,600),rng.uniform(3,100),rng.uniform(-0.8,0.8)
    c,s=math.cos(ang),math.sin(ang)
    p=np.array([[-bw/2,-bh/2],[bw/2,-bh/2],[-bw/2,bh/2],[bw/2,bh/2]])@np.array([[c,s],[-s,c]])+[cx,cy]
    src=p.astype(np.float32); dst=np.array([[0,0],[bw-1,0],[0,bh-1],[bw-1,bh-1]],np.float32)
    T=cv2.getPerspectiveTransform(src,dst); R=gpt_f32prod(src,dst)
    n+=1; ok+=int(np.array_equal(T,R))
print("float32-product hypothesis:",ok,"of",n)
