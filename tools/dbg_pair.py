import sys, os, torch
sys.path.insert(0, '/root/repo' if os.path.isdir('/root/repo/versband_b200') else '.')
import torch.nn.functional as F
from versband_b200 import ops
from versband_b200.dit import pack_conv_weight
torch.set_grad_enabled(False)
DEV='cuda'; SLOPE=0.1
def lrelu(x,s=SLOPE): return torch.where(x>0,x,x*s)
C,k,d,B,T=64,3,1,1,60000
g=torch.Generator(device=DEV).manual_seed(1)
x=torch.randn(B,T,C,device=DEV,generator=g); y=lrelu(x).half()
w1=(torch.randn(C,C,k,device=DEV,generator=g)/(C*k)**0.5).half(); w2=(torch.randn(C,C,k,device=DEV,generator=g)/(C*k)**0.5).half()
b1=torch.randn(C,device=DEV,generator=g)*0.1; b2=torch.randn(C,device=DEV,generator=g)*0.1
yf=y.float(); xr=torch.where(yf<0,yf/SLOPE,yf).permute(0,2,1)
t1=lrelu(F.conv1d(yf.permute(0,2,1),w1.float(),b1,dilation=d,padding=(k-1)//2*d)).half().float()
ref=lrelu(xr+F.conv1d(t1,w2.float(),b2,padding=(k-1)//2)).permute(0,2,1)
R=128-(k-1)
for rep in range(3):
    out=torch.zeros(B,T,C,device=DEV,dtype=torch.float16)
    ops.resblock_pair(y,out,pack_conv_weight(w1.float(),dtype=torch.float16),pack_conv_weight(w2.float(),dtype=torch.float16),b1,b2,B=B,T=T,C=C,k=k,dil=d,slope=SLOPE)
    torch.cuda.synchronize()
    diff=(out.float()-ref).abs(); bad=(diff>2e-3*float(ref.abs().max()))
    rows=bad.any(-1)[0].nonzero().flatten()
    print("rep",rep,"bad rows",rows.numel())
    for r in rows[:12].tolist():
        ch=bad[0,r].nonzero().flatten().tolist()
        print("  row",r,"tile",r//R,"local",r%R,"cta",(r//R)%148,"i",(r//R)//148,"bad ch",len(ch),ch[:10])
# hypothesis test: the bad elements used the residual of the tile that reuses the same smem stage (2 CTA-rounds later)
xr_t = xr.permute(0, 2, 1)  # [B,T,C]
pre = torch.where(ref < 0, ref / SLOPE, ref)   # x' = xr + conv
conv = pre - xr_t
shift = 2 * 148 * R
cand = torch.zeros_like(ref)
cand[:, :T - shift] = lrelu(conv[:, :T - shift] + xr_t[:, shift:])
cand0 = lrelu(conv)  # residual = 0
m = bad
if m.any():
    o = out.float()
    print("bad elems", int(m.sum()), " match shifted-residual:", int(((o - cand).abs()[m] < 2e-3).sum()),
          " match zero-residual:", int(((o - cand0).abs()[m] < 2e-3).sum()))
    rows_i = (rows // R) // 148
    print("tile-in-CTA index histogram of bad rows:", torch.bincount(rows_i).tolist())
    print("local-row min/max", int((rows % R).min()), int((rows % R).max()))
