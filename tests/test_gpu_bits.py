"""Bit-packed rasters straight into the GEMMs (SNNK_F_INPUT_BITS, csrc/gemm_bits.cuh).

k_proj_bits / k_wgrad_bits expand raster words inside their shared-memory tiles.  Their products are exact (spikes are
{0,1}; W_in as two scaled fp16 planes, gI as two tf32 planes), so -- like the fp32-fed tensor-core GEMMs -- they differ
from the oracle's fp32 sums by accumulation order only: input current within 1e-5 relative (2e-6 of max against an
fp64 product), rasters >= 99.99 % identical with every fork explained by a near-tie, gradients within 1e-4.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import OracleCfg
from _util import elementwise_err, rel_err, unexplained_forks

pytestmark = pytest.mark.gpu

from snnimageclassification_b200 import LayerType, SNN, SpikeFuncType, ToSpikes  # noqa: E402
from snnimageclassification_b200.modules import functional as F_  # noqa: E402

DEV = torch.device("cuda:0")


def npy(t):
	return None if t is None else t.detach().cpu().numpy()


def pack(x: torch.Tensor) -> torch.Tensor:
	"""(B,T,N) {0,1} -> (B,T,ceil(N/32)) int32, bit l of word w = feature 32w+l (the SNNK_BITS format), tagged."""
	B, T, N = x.shape
	W = (N + 31) // 32
	xp = np.zeros((B, T, W * 32), dtype=np.uint8)
	xp[..., :N] = npy(x).astype(np.uint8)
	words = np.packbits(xp.reshape(B, T, W, 32), axis=-1, bitorder="little").view(np.uint32).reshape(B, T, W)
	return F_.mark_bits(torch.from_numpy(words.view(np.int32)).to(DEV), N)


def _setup(B, T, N, H, O, rec, layer, density, seed=0, w_scale=None):
	g = torch.Generator().manual_seed(seed)
	theta = 0.03 if layer else 1.0
	ws = theta if w_scale is None else w_scale
	d = dict(
		x=F_.mark_binary((torch.rand(B, T, N, generator=g) < density).float().to(DEV)),
		W_in=(torch.randn(N, H, generator=g) * ws).to(DEV),
		W_rec=(torch.randn(H, H, generator=g) * theta).to(DEV) if rec else None,
		mask=(1 - torch.eye(H)).to(DEV) if rec else None,
		W_out=torch.randn(H, O, generator=g).to(DEV), b_out=(torch.randn(O, generator=g) * 0.1).to(DEV),
		beta=torch.tensor([1.6], device=DEV) if layer else None,
		labels=torch.randint(0, O, (B,), generator=g).to(DEV))
	consts = lambda tc: F_.LayerConsts(layer, 0, rec, float(np.float32(np.exp(-1 / 20))),  # noqa: E731
		float(np.float32(np.exp(-1 / 200))), theta, 0.3 if layer else 1.0, float(np.float32(np.exp(-1 / 10))), tensor_core=tc)
	return d, consts


def _fwd(d, c, x):
	return F_.run_forward(c, x, d["W_in"], d["W_rec"], d["mask"], d["beta"], d["W_out"], d["b_out"])


def _bwd(d, c, f, x, **kw):
	return F_.run_backward(c, x, d["W_rec"], d["mask"], d["beta"], d["W_out"], f["V"], f["a"], f["zbits"], Z=f["Z"], **kw)


GEOMS = [
	# B, T, N, H, rec, layer, density
	(64, 100, 784, 128, True, 1, 0.1),      # headline geometry: 25 raster words per row, the last one half used
	(64, 100, 784, 64, False, 1, 0.4),      # c3 width
	(33, 100, 784, 32, True, 0, 0.1),       # LIF, ragged row count (3300 rows: partial second row tile)
	(5, 7, 20, 32, True, 1, 0.3),           # one word per row, one k-block, 35 rows
	(3, 33, 100, 64, True, 1, 0.3),         # T not a multiple of the 32-step k-block, 4 words of which the last is partial
	(130, 1, 64, 128, True, 1, 0.2),        # T = 1
	(6, 40, 256, 256, True, 1, 0.2),        # wide layer: two n-tiles, three feature tiles + two spike tiles in K4
]


@pytest.mark.parametrize("B,T,N,H,rec,layer,density", GEOMS)
def test_projection_from_bits(B, T, N, H, rec, layer, density):
	d, consts = _setup(B, T, N, H, 10, rec, layer, density, seed=B + N)
	xb = pack(d["x"])
	f_bits = _fwd(d, consts(True), xb)
	f_simt = _fwd(d, consts(False), d["x"])
	I64 = npy(d["x"]).astype(np.float64).reshape(B * T, N) @ npy(d["W_in"]).astype(np.float64)
	got = npy(f_bits["I_in"]).reshape(B * T, H)
	assert np.abs(got - I64).max() <= 2e-6 * np.abs(I64).max(), np.abs(got - I64).max() / np.abs(I64).max()
	assert rel_err(got, npy(f_simt["I_in"]).reshape(B * T, H)) <= 1e-5
	# rasters: identical except for samples that fork at a near-tie
	thr = consts(True).theta + (1.6 * npy(f_simt["a"]) if layer else 0.0)
	forked, unexplained = unexplained_forks(npy(f_bits["Z"]), npy(f_simt["Z"]), npy(f_simt["V"]), thr)
	assert unexplained == 0
	assert forked <= max(1, B // 20)


def test_projection_from_bits_wide_dynamic_range():
	"""Columns of very different magnitude (per-column power-of-two scaling) and tiny / huge weights."""
	B, T, N, H = 16, 10, 128, 128
	d, consts = _setup(B, T, N, H, 10, False, 0, 0.3, seed=3)
	scale = torch.logspace(-12, 6, H, device=DEV)
	d["W_in"] = d["W_in"] * scale[None, :]
	d["W_in"][:, 5] = 0.0                                  # an all-zero column
	d["W_in"][::7, 9] *= 1e-6                              # small elements inside a large column
	f_bits = _fwd(d, consts(True), pack(d["x"]))
	I64 = npy(d["x"]).astype(np.float64).reshape(B * T, N) @ npy(d["W_in"]).astype(np.float64)
	got = npy(f_bits["I_in"]).reshape(B * T, H).astype(np.float64)
	colmax = np.abs(I64).max(axis=0)
	assert np.all(np.abs(got - I64).max(axis=0) <= 2e-6 * np.maximum(colmax, 1e-300))
	assert np.all(got[:, 5] == 0.0)


@pytest.mark.parametrize("B,T,N,H,rec,layer,density", GEOMS)
def test_weight_gradients_from_bits(B, T, N, H, rec, layer, density):
	"""K4 from raster words against the fp32-fed kernels on the SAME traces and seeds, and against the oracle's sweep."""
	d, consts = _setup(B, T, N, H, 10, rec, layer, density, seed=2 * B + N)
	c = consts(True)
	xb = pack(d["x"])
	f = _fwd(d, c, xb)
	loss, logp, gl = F_.run_head_nll(f["logits"], d["labels"])
	g_bits = _bwd(d, c, f, xb, g_logits=gl, tstar=f["tstar"])
	g_f32 = _bwd(d, consts(False), f, d["x"], g_logits=gl, tstar=f["tstar"])
	for k in ("dW_in", "dW_rec", "dW_out", "db"):
		if g_f32[k] is None:
			continue
		assert rel_err(npy(g_bits[k]), npy(g_f32[k])) <= 1e-5, (k, rel_err(npy(g_bits[k]), npy(g_f32[k])))
		assert elementwise_err(npy(g_bits[k]), npy(g_f32[k])) <= 1e-3, k
	if rec:
		assert np.all(np.diag(npy(g_bits["dW_rec"])) == 0.0)
	if H <= 128:
		cfg = OracleCfg(B, T, N, H, 10, layer, 0, int(rec), alpha=c.alpha, rho=c.rho, theta=c.theta, gamma=c.gamma,
			kappa=c.kappa, beta=1.6)
		hg = oracle.head(npy(f["y"]), npy(d["labels"]))
		W_rec = npy(d["W_rec"]) if rec else np.zeros((H, H), np.float32)
		mask = npy(d["mask"]) if rec else np.zeros((H, H), np.float32)
		gref = oracle.backward(cfg, npy(d["x"]), W_rec, mask, npy(d["W_out"]), npy(f["V"]),
			npy(f["a"]) if layer else np.zeros_like(npy(f["V"])), npy(f["Z"]), hg["g_y"])
		for k in ("dW_in", "dW_rec", "dW_out", "db"):
			if g_bits[k] is None:
				continue
			assert rel_err(npy(g_bits[k]), gref[k]) <= 1e-4, (k, rel_err(npy(g_bits[k]), gref[k]))


def test_headline_geometry_from_bits_vs_oracle():
	"""ALIF 784-128-10 recurrent, B = 256, T = 100 fed with the packed output of the production encoder."""
	B, T, N, H, O = 256, 100, 784, 128, 10
	d, consts = _setup(B, T, N, H, O, True, 1, 0.1, seed=11)
	g = torch.Generator().manual_seed(5)
	img = (torch.randint(1, 256, (B, N), generator=g).float() / 255.0) * (torch.rand(B, N, generator=g) < 0.19)
	enc = ToSpikes(T, use_periods=True)
	xb = F_.mark_bits(enc.encode_batch_bits(img.to(DEV)), N)
	x = enc.encode_batch(img.to(DEV), frame_runs=False)
	from snnimageclassification_b200.datasets.datasets import unpack_raster
	assert torch.equal(unpack_raster(xb, N), x)
	c = consts(True)
	cfg = OracleCfg(B, T, N, H, O, 1, 0, 1, alpha=c.alpha, rho=c.rho, theta=c.theta, gamma=c.gamma, kappa=c.kappa, beta=1.6)
	f = _fwd(d, c, xb)
	ref = oracle.forward(cfg, npy(x), npy(d["W_in"]), npy(d["W_rec"]), npy(d["mask"]), npy(d["W_out"]), npy(d["b_out"]))
	assert rel_err(npy(f["I_in"]), ref["I_in"]) <= 1e-5
	forked, unexplained = unexplained_forks(npy(f["Z"]), ref["Z"], ref["V"], c.theta + 1.6 * ref["a"])
	assert unexplained == 0
	same = (npy(f["Z"]) == ref["Z"]).mean()
	assert same >= 0.9999 or forked <= 1, (same, forked)
	labels = npy(d["labels"])
	hg = oracle.head(npy(f["y"]), labels)
	loss, logp, gl = F_.run_head_nll(f["logits"], d["labels"])
	gr = _bwd(d, c, f, xb, g_logits=gl, tstar=f["tstar"])
	gref = oracle.backward(cfg, npy(x), npy(d["W_rec"]), npy(d["mask"]), npy(d["W_out"]), npy(f["V"]), npy(f["a"]),
		npy(f["Z"]), hg["g_y"])
	for k in ("dW_in", "dW_rec", "dW_out", "db"):
		assert rel_err(npy(gr[k]), gref[k]) <= 1e-4, (k, rel_err(npy(gr[k]), gref[k]))
		assert elementwise_err(npy(gr[k]), gref[k]) <= 1e-3, k


def test_bits_flag_needs_tensor_core_mode():
	d, consts = _setup(4, 5, 64, 32, 10, True, 1, 0.3)
	with pytest.raises(RuntimeError, match="not implemented"):
		_fwd(d, consts(False), pack(d["x"]))


@pytest.mark.parametrize("H,layer", [(128, LayerType.ALIF), (64, LayerType.LIF), (256, LayerType.ALIF)])
def test_snn_module_takes_packed_rasters(H, layer, monkeypatch):
	"""SNN keeps an int32 packed raster packed (training step and no-trace inference) and gives what the unpacked
	raster gives; SNNK_PACKED_GEMM=0 restores the unpack-first behaviour."""
	B, T, N = 48, 30, 784
	torch.manual_seed(0)
	net = SNN(N, 10, H, use_recurrent_connection=True, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=layer, device=DEV, **({"learn_beta": True} if layer == LayerType.ALIF else {}))
	g = torch.Generator().manual_seed(1)
	img = (torch.randint(1, 256, (B, N), generator=g).float() / 255.0) * (torch.rand(B, N, generator=g) < 0.19)
	lab = torch.randint(0, 10, (B,), generator=g).to(DEV)
	enc = ToSpikes(T, use_periods=True, tau=20.0)          # the reference's test regime: latencies spread over T
	bits = enc.encode_batch_bits(img.to(DEV))
	x = enc.encode_batch(img.to(DEV), frame_runs=False)
	net.train()

	def grads(inp):
		net.zero_grad()
		loss = net.batch_loss(inp, lab)
		loss.backward()
		return loss.item(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
	l_bits, g_bits = grads(bits)
	l_x, g_x = grads(x)
	monkeypatch.setenv("SNNK_PACKED_GEMM", "0")
	l_unp, g_unp = grads(bits)
	monkeypatch.delenv("SNNK_PACKED_GEMM")
	assert abs(l_bits - l_x) <= 1e-4 * abs(l_x) and abs(l_unp - l_x) <= 1e-6 * abs(l_x)
	for n in g_x:
		assert rel_err(npy(g_bits[n]), npy(g_x[n])) <= 1e-4, (n, rel_err(npy(g_bits[n]), npy(g_x[n])))
	net.eval()
	with torch.no_grad():
		lo_bits = net.get_prediction_logits(bits, re_outputs_trace=False, re_hidden_states=False)
		lo_x = net.get_prediction_logits(x, re_outputs_trace=False, re_hidden_states=False)
		y_bits, hs = net(bits)
	assert rel_err(npy(lo_bits), npy(lo_x)) <= 1e-4
	assert y_bits.shape == (B, T, 10) and hs["input"][0].shape == (B, T, H)
