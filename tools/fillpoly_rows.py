# brute-force check: per-row interval form of on_line() == per-pixel predicate
def on_line(px,py,x1,y1,x2,y2):
    dx=x2-x1; dy=y2-y1
    if dx<0:
        x1,x2=x2,x1; y1,y2=y2,y1; dx=-dx; dy=-dy
    ystep=1
    if dy<0: dy=-dy; ystep=-1
    if dy>dx:
        j=(py-y1)*ystep
        if j<0 or j>dy: return False
        T=2*dx*j-dy
        k=0 if T<=0 else (T+2*dy-1)//(2*dy)
        return px==x1+k
    i=px-x1
    if i<0 or i>dx: return False
    if dx==0: return py==y1
    T=2*dy*i-dx
    k=0 if T<=0 else (T+2*dx-1)//(2*dx)
    return py==y1+ystep*k
def row_interval(py,x1,y1,x2,y2):
    """returns (lo,hi) inclusive px interval, empty if lo>hi"""
    dx=x2-x1; dy=y2-y1
    if dx<0:
        x1,x2=x2,x1; y1,y2=y2,y1; dx=-dx; dy=-dy
    ystep=1
    if dy<0: dy=-dy; ystep=-1
    if dy>dx:
        j=(py-y1)*ystep
        if j<0 or j>dy: return (1,0)
        T=2*dx*j-dy
        k=0 if T<=0 else (T+2*dy-1)//(2*dy)
        return (x1+k,x1+k)
    if dx==0:
        return (x1,x1) if py==y1 else (1,0)
    kk=(py-y1)*ystep
    if kk<0 or kk>dy: return (1,0)
    if dy==0: return (x1,x1+dx)
    lo = 0 if kk==0 else (dx*(2*kk-1))//(2*dy)+1
    hi = (dx*(2*kk+1))//(2*dy)
    if hi>dx: hi=dx
    return (x1+lo,x1+hi)
import itertools
bad=0
R=14
for x1,y1,x2,y2 in itertools.product(range(0,R,1),range(0,R),range(0,R),range(0,R)):
    for py in range(-1,R+1):
        lo,hi=row_interval(py,x1,y1,x2,y2)
        for px in range(-1,R+1):
            a=on_line(px,py,x1,y1,x2,y2); b=lo<=px<=hi
            if a!=b:
                bad+=1
                if bad<10: print(x1,y1,x2,y2,px,py,a,b,lo,hi)
print("bad",bad)
import random
random.seed(1)
for _ in range(20000):
    x1,y1,x2,y2=[random.randint(0,400) for _ in range(4)]
    for py in random.sample(range(-2,403),12):
        lo,hi=row_interval(py,x1,y1,x2,y2)
        for px in range(-1,402):
            if on_line(px,py,x1,y1,x2,y2)!=(lo<=px<=hi): bad+=1
print("bad",bad)
