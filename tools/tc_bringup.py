"""Bring-up check of the tcgen05 GEMMs against the fp32 SIMT kernels (run on the GPU box).

    python tools/tc_bringup.py proj|wgrad [H]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from snnimageclassification_b200.modules import functional as F_

which = sys.argv[1]
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
B, T, N, O = (int(v) for v in (sys.argv[3:7] if len(sys.argv) > 6 else (16, 100, 784, 10)))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = (torch.rand(B, T, N, generator=g) < 0.15).float().to(dev)
W_in = (torch.randn(N, H, generator=g) * 0.03).to(dev)
W_rec = (torch.randn(H, H, generator=g) * 0.03).to(dev)
mask = (1 - torch.eye(H)).to(dev)
W_out = torch.randn(H, O, generator=g).to(dev)
b_out = torch.zeros(O).to(dev)
beta = torch.tensor([1.6], device=dev)
labels = torch.randint(0, O, (B,), generator=g).to(dev)


def consts(tc):
	return F_.LayerConsts(1, 0, True, float(np.exp(-1 / 20)), float(np.exp(-1 / 200)), 0.03, 0.3, float(np.exp(-1 / 10)), tensor_core=tc)


def rel(a, b):
	return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


print("device", torch.cuda.get_device_name(0), "H", H, "B,T,N", B, T, N, flush=True)
ref = F_.run_forward(consts(False), x, W_in, W_rec, mask, beta, W_out, b_out)
torch.cuda.synchronize()
print("simt forward ok", flush=True)
if which == "proj":
	out = F_.run_forward(consts(True), x, W_in, W_rec, mask, beta, W_out, b_out)
	torch.cuda.synchronize()
	print("tc forward returned", flush=True)
	I_ref, I_tc = ref["I_in"].double(), out["I_in"].double()
	exact = x.double().reshape(-1, N) @ W_in.double()
	print("I_in rel err tc vs simt :", rel(I_tc, I_ref))
	print("I_in rel err tc vs fp64 :", rel(I_tc.reshape(-1, H), exact), " simt vs fp64:", rel(I_ref.reshape(-1, H), exact))
	d = (I_tc - I_ref).abs()
	print("max abs diff", float(d.max()), "mean abs diff", float(d.mean()), "mean |I|", float(I_ref.abs().mean()))
	print("bit-identical fraction", float((out["I_in"] == ref["I_in"]).float().mean()))
	print("raster identical fraction", float((out["Z"] == ref["Z"]).float().mean()))
	# non-exact input must fall back and be bit-identical
	x2 = x * 0.3
	r2 = F_.run_forward(consts(False), x2, W_in, W_rec, mask, beta, W_out, b_out)
	o2 = F_.run_forward(consts(True), x2, W_in, W_rec, mask, beta, W_out, b_out)
	torch.cuda.synchronize()
	print("fallback bit-identical:", bool(torch.equal(r2["I_in"], o2["I_in"])))
else:
	loss, logp, gl = F_.run_head_nll(ref["logits"], labels)
	kw = dict(g_logits=gl, tstar=ref["tstar"], Z=ref["Z"])
	g0 = F_.run_backward(consts(False), x, W_rec, mask, beta, W_out, ref["V"], ref["a"], ref["zbits"], **kw)
	torch.cuda.synchronize()
	print("simt backward ok", flush=True)
	g1 = F_.run_backward(consts(True), x, W_rec, mask, beta, W_out, ref["V"], ref["a"], ref["zbits"], **kw)
	torch.cuda.synchronize()
	print("tc backward returned", flush=True)
	for k in ("gI", "dW_in", "dW_rec", "dW_out", "db"):
		print(k, "rel err tc vs simt:", rel(g1[k].double(), g0[k].double()))
	x2 = x * 0.3
	g2 = F_.run_backward(consts(False), x2, W_rec, mask, beta, W_out, ref["V"], ref["a"], ref["zbits"], **kw)
	g3 = F_.run_backward(consts(True), x2, W_rec, mask, beta, W_out, ref["V"], ref["a"], ref["zbits"], **kw)
	torch.cuda.synchronize()
	print("fallback dW_in rel err:", rel(g3["dW_in"].double(), g2["dW_in"].double()))
print("done", flush=True)
