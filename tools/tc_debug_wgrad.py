import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
from snnimageclassification_b200.modules import functional as F_
from snnimageclassification_b200 import _cabi
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B, T, N = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (4, 32, 256)
O = 10
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = (torch.rand(B, T, N, generator=g) < 0.3).float().to(dev)
W_rec = (torch.randn(H, H, generator=g) * 0.03).to(dev); mask=(1-torch.eye(H)).to(dev)
W_out = torch.randn(H, O, generator=g).to(dev); beta = torch.tensor([1.6], device=dev)
V = torch.randn(B,T,H, generator=g).to(dev)*0.05; a = torch.rand(B,T,H, generator=g).to(dev)
Z = (torch.rand(B,T,H, generator=g) < 0.4).float().to(dev)
zb = (Z.reshape(B,T,H//32,32).to(torch.int64) << torch.arange(32, device=dev)).sum(-1)
zbits = torch.where(zb >= 2**31, zb - 2**32, zb).to(torch.int32).contiguous()
g_y = torch.randn(B,T,O, generator=g).to(dev)
def consts(tc): return F_.LayerConsts(1,0,True,0.95,0.99,0.03,0.3,0.9,tensor_core=tc)
g0 = F_.run_backward(consts(False), x, W_rec, mask, beta, W_out, V, a, zbits, g_y=g_y, Z=Z)
g1 = F_.run_backward(consts(True), x, W_rec, mask, beta, W_out, V, a, zbits, g_y=g_y, Z=Z)
torch.cuda.synchronize()
gI = g0["gI"]().double().reshape(B*T, H)
exp_in = x.double().reshape(B*T, N).t() @ gI
Zs = torch.cat([torch.zeros(B,1,H,device=dev), Z[:,:-1]],1).double().reshape(B*T,H)
exp_rec = (Zs.t() @ gI) * mask.double()
for name, got, exp in (("dW_in", g1["dW_in"], exp_in), ("dW_rec", g1["dW_rec"], exp_rec), ("simt dW_in", g0["dW_in"], exp_in)):
	got = got.double()
	print(name, "max|got|", float(got.abs().max()), "max|exp|", float(exp.abs().max()), "frac zero", float((got==0).double().mean()),
		"err", float((got-exp).abs().max()/exp.abs().max()))
	if name != "simt dW_in":
		# pattern probes
		print("   corr with exp:", float((got*exp).sum()/ (got.norm()*exp.norm()+1e-30)))
		print("   got[:4,:6]", got[:4,:6].cpu().numpy().round(4).tolist())
		print("   exp[:4,:6]", exp[:4,:6].cpu().numpy().round(4).tolist())
		nzr = (got.abs().sum(1) > 0).nonzero().flatten()[:20].tolist(); nzc = (got.abs().sum(0) > 0).nonzero().flatten()[:20].tolist()
		print("   nonzero rows", nzr, "cols", nzc)

gI32 = g0["gI"]().reshape(B*T, H)
hi = (gI32.view(torch.int32) & -8192).view(torch.float32).double()
lo = gI.double() - hi
X = x.double().reshape(B*T, N)
got = g1["dW_in"].double()
for name, e in (("hi only", X.t() @ hi), ("hi+lo", X.t() @ (hi+lo)), ("hi+2lo", X.t() @ (hi+2*lo)), ("2hi", X.t() @ (2*hi))):
	print(name, float((got - e).abs().max() / e.abs().max()))
err = (got - exp_in).abs()
print("rows with err>1e-3*max:", (err.max(1)[0] > 1e-3*exp_in.abs().max()).nonzero().flatten()[:40].tolist())
print("cols with err>1e-3*max:", (err.max(0)[0] > 1e-3*exp_in.abs().max()).nonzero().flatten()[:40].tolist())
