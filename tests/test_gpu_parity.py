"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's golden fixtures.

Bars (BASELINE.json north_star): to_spikes bit-exact; fp32 mode V/a within 1e-5 relative and spike rasters
>= 99.99 % identical; loss and gradients within 1e-4 relative.  With the fp32 SIMT projection the forward is in
fact bit-identical to the C oracle (same summation orders), which the tests assert.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import OracleCfg
from oracle.torch_port import TorchPortSNN
from _util import dynamics_case, first_divergence, load, rel_err, unpack_bits

pytestmark = pytest.mark.gpu

from snnimageclassification_b200 import _cabi  # noqa: E402
from snnimageclassification_b200 import LayerType, SNN, SpikeFuncType, ToSpikes  # noqa: E402
from snnimageclassification_b200.modules import functional as F_  # noqa: E402
from snnimageclassification_b200.modules.spike_funcs import HeavisidePhiApprox, HeavisideSigmoidApprox  # noqa: E402

DEV = torch.device("cuda:0")


def cu(a, dtype=torch.float32):
	return None if a is None else torch.as_tensor(np.ascontiguousarray(a)).to(DEV, dtype)


def npy(t):
	return None if t is None else t.detach().cpu().numpy()


def test_native_library_loaded_and_device_supported():
	assert _cabi.lib().snnk_abi_version() == 8
	_cabi.require_b200(DEV)


# ---- encoder: bit-exact ------------------------------------------------------------------------------------------
def test_encoder_reference_known_answers():
	"""The reference's own tests (test/test_to_spikes.py), run against the CUDA encoder."""
	tr = ToSpikes(100, 100, tau=20.0, thr=0.2, epsilon=1e-7)
	assert np.all(tr.pixels_to_firing_periods(np.array([0.0])) == tr.n_steps)                       # :9-13
	pix = np.array([0.82352941, 0.82745098, 0.83529412, 0.8745098, 0.8627451, 0.95294118, 0.79215686, 0., 0., 0.])
	assert np.array_equal(tr.pixels_to_firing_periods(pix), [5, 5, 5, 5, 5, 4, 5, 100, 100, 100])  # :15-20
	tr = ToSpikes(10, 10, tau=20.0, thr=0.2, epsilon=1e-7)
	pix = np.array([0.8627451, 0.90980392, 0.96470588, 0., 0.01176471, 0.79215686, 0.89411765, 0.87843137,
		0.86666667, 0.82745098, 0.82745098, 0.83921569])
	exp = np.zeros((10, 12))
	exp[[4, 4, 5, 5, 5, 5, 5, 5, 5, 5], [1, 2, 0, 5, 6, 7, 8, 9, 10, 11]] = 1
	out = tr(pix)
	assert out.dtype == torch.float64 and tuple(out.shape) == (10, 12)
	assert np.array_equal(out.numpy(), exp)                                                         # :38-50
	ft = np.array([5, 4, 4, 10, 10, 5, 5, 5, 5, 5, 5, 5])
	assert np.array_equal(tr.firing_times_to_spikes(ft), exp)                                       # :52-60
	tr = ToSpikes(5, 5)
	exp = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 0, 0], [1, 1, 1]])
	assert np.array_equal(tr.firing_periods_to_spikes(np.array([1, 2, tr.n_steps + 1])), exp)       # :62-73


def test_encoder_golden_image():
	z = load("encoder_golden.npz")                                                                   # :75-83
	x = z["real_x_f64"]
	spikes = unpack_bits(z["real_spikes_bits"], tuple(z["real_spikes_shape"]))
	tr = ToSpikes(100, 100, tau=20.0, thr=0.2, epsilon=1e-7)
	assert np.array_equal(tr.pixels_to_firing_periods(x), z["real_periods"])
	assert np.array_equal(tr(torch.from_numpy(x)).numpy(), spikes)


def test_encoder_against_reference_fixtures():
	z = load("encoder_golden.npz")
	for key in z["cases"]:
		x = z[f"{key}_x"]
		parts = key.split("_")
		tau = float(parts[2][3:]); periodic = bool(int(parts[3][1:])); n = int(parts[4][1:])
		tr = ToSpikes(n, n, tau=tau, thr=0.2, use_periods=periodic, epsilon=1e-7)
		assert np.array_equal(tr.pixels_to_firing_periods(x), z[f"{key}_periods"]), key
		for dt in (torch.float32, torch.float64, torch.uint8):
			ras = tr.encode_batch(torch.from_numpy(x), out_dtype=dt)
			assert ras.dtype == dt and ras.is_cuda
			assert np.array_equal(npy(ras).astype(np.uint8), unpack_bits(z[f"{key}_bits"], tuple(ras.shape))), key


@pytest.mark.parametrize("periodic", [False, True])
@pytest.mark.parametrize("tau", [20.0, 0.02])
def test_encoder_full_size_vs_oracle(periodic, tau):
	"""BASELINE-sized batch (256 x 784, T = 100) of k/255 images: bit-exact against the oracle."""
	g = torch.Generator().manual_seed(3)
	img = (torch.randint(0, 256, (256, 784), generator=g).float() / 255.0) * (torch.rand(256, 784, generator=g) < 0.3)
	tr = ToSpikes(100, use_periods=periodic, tau=tau)
	got = tr.encode_batch(img.to(DEV), out_dtype=torch.uint8)
	want = oracle.encode(img.numpy(), 100, None, tau=tau, thr=0.2, periodic=periodic, eps=1e-7)
	assert np.array_equal(npy(got), want)
	if periodic and tau == 0.02:      # production regime: last frame is all ones
		assert npy(got)[:, -1].all() and not npy(got)[:, 0].any()


def test_encoder_ragged_and_empty():
	tr = ToSpikes(7, use_periods=True, tau=20.0)
	assert tuple(tr.encode_batch(torch.zeros((0, 5))).shape) == (0, 7, 5)
	x = torch.rand(3, 1, generator=torch.Generator().manual_seed(0))
	assert np.array_equal(npy(tr.encode_batch(x, out_dtype=torch.uint8)),
		oracle.encode(x.numpy(), 7, None, tau=20.0, thr=0.2, periodic=True, eps=1e-7))


# ---- dynamics ------------------------------------------------------------------------------------------------------
def _cfg(d):
	B, T, N, H, O = (int(v) for v in d["dims"])
	alif, phi, rec, _ = (int(v) for v in d["flags"])
	al, rho, th, ga, ka, be = (float(v) for v in d["scalars"])
	return OracleCfg(B, T, N, H, O, layer_type=alif, surrogate=phi, recurrent=rec, alpha=al, rho=rho, theta=th,
		gamma=ga, kappa=ka, beta=be)


def _consts(cfg):
	return F_.LayerConsts(cfg.layer_type, cfg.surrogate, bool(cfg.recurrent), cfg.alpha, cfg.rho, cfg.theta,
		cfg.gamma, cfg.kappa)


def _gpu_forward(cfg, x, W_in, W_rec, mask, W_out, b_out, **kw):
	beta = torch.tensor([cfg.beta], dtype=torch.float32, device=DEV) if cfg.layer_type == 1 else None
	return F_.run_forward(_consts(cfg), cu(x), cu(W_in), cu(W_rec), cu(mask), beta, cu(W_out), cu(b_out), **kw), beta


NAMES = [str(n) for n in load("dynamics_golden.npz")["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_forward_backward_vs_reference_and_oracle(name):
	d = dynamics_case(load("dynamics_golden.npz"), name)
	cfg = _cfg(d)
	x = d["x"].astype(np.float32)
	out, beta = _gpu_forward(cfg, x, d["W_in"], d.get("W_rec"), d.get("rec_mask"), d["W_out"], d["b_out"])
	f = oracle.forward(cfg, x, d["W_in"], d.get("W_rec"), d.get("rec_mask"), d["W_out"], d["b_out"])
	# (1) bit-identical to the C oracle
	assert np.array_equal(npy(out["I_in"]), f["I_in"]), "projection differs from the oracle"
	assert np.array_equal(npy(out["Z"]), f["Z"])
	assert np.array_equal(npy(out["V"]), f["V"])
	if cfg.layer_type == 1:
		assert np.array_equal(npy(out["a"]), f["a"])
	assert np.array_equal(npy(out["y"]), f["y"])
	# (2) against the reference's own outputs: rasters >= 99.99 %, state 1e-5 relative
	Zref = d["Z"].astype(np.float32)
	assert (npy(out["Z"]) == Zref).mean() >= 0.9999
	assert (first_divergence(npy(out["Z"]), Zref) == cfg.T).all()
	assert rel_err(npy(out["V"]), d["V"]) <= 1e-5
	if cfg.layer_type == 1:
		assert rel_err(npy(out["a"]), d["a"]) <= 1e-5
	assert rel_err(npy(out["y"]), d["y"]) <= 1e-5
	# bit-packed raster and fused max-over-time
	zb = npy(out["zbits"]).view(np.uint32)
	Z = npy(out["Z"])
	bits = ((zb[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(cfg.B, cfg.T, cfg.H)
	assert np.array_equal(bits, Z.astype(np.uint32))
	h = oracle.head(f["y"], d["labels"])
	assert np.array_equal(npy(out["logits"]), h["logits"]) and np.array_equal(npy(out["tstar"]), h["tstar"])
	# (3) head + BPTT
	loss, logp, g_logits = F_.run_head_nll(out["logits"], cu(d["labels"], torch.int64))
	assert rel_err(npy(logp), d["logp"]) <= 1e-5
	assert abs(loss.item() - float(d["loss"])) <= 1e-5 * abs(float(d["loss"]))
	gref = oracle.backward(cfg, x, d.get("W_rec"), d.get("rec_mask"), d["W_out"], f["V"], f["a"], f["Z"], h["g_y"])
	for mode in ("sparse", "dense"):
		kw = dict(g_logits=g_logits, tstar=out["tstar"]) if mode == "sparse" else dict(g_y=cu(h["g_y"]))
		g = F_.run_backward(_consts(cfg), cu(x), cu(d.get("W_rec")), cu(d.get("rec_mask")), beta, cu(d["W_out"]),
			out["V"], out["a"], out["zbits"], **kw)
		assert rel_err(npy(g["gI"]()), gref["gI"]) <= 1e-5, mode
		for k in ("dW_in", "dW_out", "db") + (("dW_rec",) if cfg.recurrent else ()):
			assert rel_err(npy(g[k]), d[k]) <= 1e-4, (mode, k)       # vs the reference's autograd
			assert rel_err(npy(g[k]), gref[k]) <= 1e-5, (mode, k)     # vs the oracle
		if cfg.recurrent:
			assert np.all(np.diag(npy(g["dW_rec"])) == 0.0)


@pytest.mark.parametrize("layer,rec,phi,H", [(1, 1, 0, 128), (0, 1, 0, 128), (1, 0, 1, 64), (0, 1, 1, 32)])
def test_baseline_sized_batch_vs_oracle(layer, rec, phi, H):
	"""B = 256, T = 100, N = 784 (BASELINE configs[0..2] geometry) against the C oracle: forward bit-exact."""
	B, T, N, O = 256, 100, 784, 10
	rng = np.random.default_rng(11 + H)
	theta = 0.03 if layer else 1.0
	W_in = (rng.standard_normal((N, H)) * theta).astype(np.float32)
	W_rec = (rng.standard_normal((H, H)) * theta).astype(np.float32) if rec else None
	mask = (1 - np.eye(H)).astype(np.float32) if rec else None
	W_out = rng.standard_normal((H, O)).astype(np.float32)
	b_out = (rng.standard_normal(O) * 0.1).astype(np.float32)
	x = (rng.random((B, T, N)) < (0.1 if layer else 0.02)).astype(np.float32)
	labels = rng.integers(0, O, B)
	cfg = OracleCfg(B, T, N, H, O, layer_type=layer, surrogate=phi, recurrent=rec, alpha=float(np.float32(np.exp(-1 / 20))),
		rho=float(np.float32(np.exp(-1 / 200))), theta=theta, gamma=0.3 if layer else 1.0,
		kappa=float(np.float32(np.exp(-1 / 10))), beta=1.6)
	out, beta = _gpu_forward(cfg, x, W_in, W_rec, mask, W_out, b_out)
	f = oracle.forward(cfg, x, W_in, W_rec, mask, W_out, b_out)
	for k in ("I_in", "V", "Z", "y") + (("a",) if layer else ()):
		assert np.array_equal(npy(out[k]), f[k]), k
	assert 0.001 < f["Z"].mean() < 0.999
	h = oracle.head(f["y"], labels)
	loss, logp, g_logits = F_.run_head_nll(out["logits"], cu(labels, torch.int64))
	assert abs(loss.item() - h["loss"]) <= 1e-5 * abs(h["loss"])
	g = F_.run_backward(_consts(cfg), cu(x), cu(W_rec), cu(mask), beta, cu(W_out), out["V"], out["a"], out["zbits"],
		g_logits=g_logits, tstar=out["tstar"])
	gref = oracle.backward(cfg, x, W_rec, mask, W_out, f["V"], f["a"], f["Z"], h["g_y"])
	assert rel_err(npy(g["gI"]()), gref["gI"]) <= 1e-5
	for k in ("dW_in", "dW_out", "db") + (("dW_rec",) if rec else ()):
		assert rel_err(npy(g[k]), gref[k]) <= 1e-4, k
	# inference mode (no traces) gives the same logits
	out2, _ = _gpu_forward(cfg, x, W_in, W_rec, mask, W_out, b_out, traces=False)
	assert out2["V"] is None and np.array_equal(npy(out2["logits"]), npy(out["logits"]))


def test_ragged_batch_sizes_and_short_sequences():
	"""B not a multiple of the rows-per-CTA, T not a multiple of the prefetch depth, T = 1."""
	rng = np.random.default_rng(5)
	for B, T, H in ((1, 1, 32), (3, 7, 64), (1031, 5, 128), (5, 3, 128)):
		N, O = 20, 10
		W_in = (rng.standard_normal((N, H)) * 0.3).astype(np.float32)
		W_rec = (rng.standard_normal((H, H)) * 0.1).astype(np.float32)
		W_out = rng.standard_normal((H, O)).astype(np.float32)
		b_out = rng.standard_normal(O).astype(np.float32)
		x = (rng.random((B, T, N)) < 0.3).astype(np.float32)
		cfg = OracleCfg(B, T, N, H, O, 1, 0, 1, alpha=0.95, rho=0.99, theta=0.3, gamma=0.3, kappa=0.9, beta=0.5)
		out, beta = _gpu_forward(cfg, x, W_in, W_rec, None, W_out, b_out)
		f = oracle.forward(cfg, x, W_in, W_rec, None, W_out, b_out)
		for k in ("V", "a", "Z", "y"):
			assert np.array_equal(npy(out[k]), f[k]), (B, T, H, k)
		labels = rng.integers(0, O, B)
		h = oracle.head(f["y"], labels)
		g = F_.run_backward(_consts(cfg), cu(x), cu(W_rec), None, beta, cu(W_out), out["V"], out["a"], out["zbits"],
			g_y=cu(h["g_y"]))
		gref = oracle.backward(cfg, x, W_rec, None, W_out, f["V"], f["a"], f["Z"], h["g_y"])
		for k in ("dW_in", "dW_rec", "dW_out", "db"):
			assert rel_err(npy(g[k]), gref[k]) <= 1e-4, (B, T, H, k)


def test_initial_state_and_hidden_trace_seeds():
	rng = np.random.default_rng(9)
	B, T, N, H, O = 4, 6, 16, 32, 10
	W_in = (rng.standard_normal((N, H)) * 0.3).astype(np.float32)
	W_rec = (rng.standard_normal((H, H)) * 0.1).astype(np.float32)
	W_out = rng.standard_normal((H, O)).astype(np.float32)
	b_out = np.zeros(O, np.float32)
	x = (rng.random((B, T, N)) < 0.3).astype(np.float32)
	V0 = rng.standard_normal((B, H)).astype(np.float32) * 0.1
	a0 = rng.random((B, H)).astype(np.float32)
	Z0 = (rng.random((B, H)) < 0.3).astype(np.float32)
	cfg = OracleCfg(B, T, N, H, O, 1, 0, 1, alpha=0.95, rho=0.99, theta=0.3, gamma=0.3, kappa=0.9, beta=0.5)
	out, beta = _gpu_forward(cfg, x, W_in, W_rec, None, W_out, b_out, state=(cu(V0), cu(a0), cu(Z0)))
	f = oracle.forward(cfg, x, W_in, W_rec, None, W_out, b_out, V0=V0, a0=a0, Z0=Z0)
	for k in ("V", "a", "Z", "y"):
		assert np.array_equal(npy(out[k]), f[k]), k
	g_y = rng.standard_normal((B, T, O)).astype(np.float32)
	g_V = rng.standard_normal((B, T, H)).astype(np.float32)
	g_Z = rng.standard_normal((B, T, H)).astype(np.float32)
	g = F_.run_backward(_consts(cfg), cu(x), cu(W_rec), None, beta, cu(W_out), out["V"], out["a"], out["zbits"],
		g_y=cu(g_y), g_V=cu(g_V), g_Z=cu(g_Z), Z0=cu(Z0))
	gref = oracle.backward(cfg, x, W_rec, None, W_out, f["V"], f["a"], f["Z"], g_y, Z0=Z0, g_Vs=g_V, g_Zs=g_Z)
	for k in ("gI", "dW_in", "dW_rec", "dW_out", "db"):
		assert rel_err(npy(g[k]() if callable(g[k]) else g[k]), gref[k]) <= 1e-5, k


# ---- the reference-facing Python API ---------------------------------------------------------------------------------
def _make_pair(layer, rec, sf, lb, N=40, H=32, O=10, T=12, seed=0):
	torch.manual_seed(seed)
	kw = dict(learn_beta=lb) if layer == LayerType.ALIF else {}
	net = SNN(N, O, H, use_recurrent_connection=rec, int_time_steps=T, spike_func=sf, hidden_layer_type=layer,
		device=DEV, **kw)
	L, R = net.layers["input"], net.layers["readout"]
	port = TorchPortSNN(N, H, O, T, layer_type=int(layer == LayerType.ALIF), surrogate=int(sf == SpikeFuncType.Phi),
		recurrent=rec, learn_beta=lb, seed=0)
	port.load(npy(L.forward_weights), npy(L.recurrent_weights) if rec else None, npy(R.forward_weights),
		npy(R.bias_weights), beta=float(L.beta) if layer == LayerType.ALIF else None)
	return net, port


@pytest.mark.parametrize("layer", [LayerType.LIF, LayerType.ALIF])
@pytest.mark.parametrize("rec", [False, True])
@pytest.mark.parametrize("sf", [SpikeFuncType.FastSigmoid, SpikeFuncType.Phi])
def test_snn_module_vs_torch_port(layer, rec, sf):
	"""SNN(...) forward / log-proba / loss / grads against the PyTorch-CPU restatement of the reference."""
	lb = layer == LayerType.ALIF
	net, port = _make_pair(layer, rec, sf, lb)
	g = torch.Generator().manual_seed(4)
	x = (torch.rand(6, 12, 40, generator=g) < (0.2 if layer == LayerType.ALIF else 0.05)).float()
	labels = torch.randint(0, 10, (6,), generator=g)
	net.train()
	logp, out, hs = net.get_prediction_log_proba(x, re_outputs_trace=True, re_hidden_states=True)
	p_logp, p_out, p_hs = port.log_proba(x)
	assert out.shape == (6, 12, 10) and len(hs["input"]) == (3 if layer == LayerType.ALIF else 2)
	assert hs["readout"][0] is out
	assert np.array_equal(npy(hs["input"][-1]), npy(p_hs["input"][-1]))        # rasters identical
	assert rel_err(npy(hs["input"][0]), npy(p_hs["input"][0])) <= 1e-5
	assert rel_err(npy(out), npy(p_out)) <= 1e-5 and rel_err(npy(logp), npy(p_logp)) <= 1e-5
	# generic autograd path with a user criterion
	crit = torch.nn.NLLLoss()
	loss_generic = torch.nn.functional.nll_loss(logp, labels.to(DEV))
	net.zero_grad()
	loss_generic.backward()
	grads_generic = [p.grad.clone() if p.grad is not None else None for p in net.parameters()]
	# fused path
	net.zero_grad()
	loss_fused = net.batch_loss(x, labels, crit)
	loss_fused.backward()
	p_loss = port.exec_batch(x, labels)
	assert abs(loss_fused.item() - p_loss) <= 1e-5 * abs(p_loss)
	assert abs(loss_generic.item() - p_loss) <= 1e-5 * abs(p_loss)
	L, R = net.layers["input"], net.layers["readout"]
	pairs = [(L.forward_weights, port.W_in), (R.forward_weights, port.W_out), (R.bias_weights, port.b_out)]
	if rec:
		pairs.append((L.recurrent_weights, port.W_rec))
	for mine, theirs in pairs:
		assert rel_err(npy(mine.grad), npy(theirs.grad)) <= 1e-4
	for gg, p in zip(grads_generic, net.parameters()):
		if p.grad is None:
			assert gg is None
		else:
			assert rel_err(npy(gg), npy(p.grad)) <= 1e-5
	if lb:
		assert L.beta.grad is None and port.beta.grad is None     # reference quirk


def test_prediction_heads_arity_and_inference_path():
	net, port = _make_pair(LayerType.ALIF, True, SpikeFuncType.FastSigmoid, True)
	x = (torch.rand(5, 12, 40, generator=torch.Generator().manual_seed(8)) < 0.2).float()
	net.eval()
	with torch.no_grad():
		lg, tr, hs = net.get_prediction_logits(x)
		lg2, tr2 = net.get_prediction_logits(x, re_hidden_states=False)
		lg3, hs3 = net.get_prediction_logits(x, re_outputs_trace=False)
		lg4 = net.get_prediction_logits(x, re_outputs_trace=False, re_hidden_states=False)
		pr, *_ = net.get_prediction_proba(x)
	assert torch.equal(lg, lg2) and torch.equal(lg, lg3) and torch.equal(lg, lg4)
	assert torch.equal(lg, tr.max(dim=1)[0])
	assert torch.allclose(pr.sum(-1), torch.ones(5, device=DEV), atol=1e-6)
	# (B, F) inputs are repeated over time, short sequences are zero padded (snn.py:159-184)
	with torch.no_grad():
		y_rep, _ = net(x[:, 0])
		y_full, _ = net(x[:, :1].repeat(1, 12, 1))
		y_pad, _ = net(x[:, :5])
		y_pad2, _ = net(torch.cat([x[:, :5], torch.zeros(5, 7, 40)], dim=1))
	assert torch.equal(y_rep, y_full) and torch.equal(y_pad, y_pad2)


def test_spike_functions_standalone():
	v = torch.linspace(-1, 3, 1001, device=DEV, requires_grad=True)
	th, ga = torch.tensor(1.0, device=DEV), torch.tensor(0.7, device=DEV)
	for fn, ref in ((HeavisideSigmoidApprox, lambda v_: 1 / (0.7 * (v_ - 1).abs() + 1) ** 2),
			(HeavisidePhiApprox, lambda v_: (0.7 / (1 + 1e-5)) * torch.clamp(1 - ((v_ - 1) / (1 + 1e-5)).abs(), min=0))):
		z = fn.apply(v, th, ga)
		assert torch.equal(z, (v >= 1.0).float())
		g, = torch.autograd.grad(z.sum(), v)
		assert torch.allclose(g, ref(v.detach()), rtol=1e-6, atol=1e-7)
	thr_t = torch.full_like(v, 0.5)
	assert torch.equal(HeavisideSigmoidApprox.apply(v.detach(), thr_t, ga), (v >= 0.5).float())


def test_layer_single_step_api():
	"""LIFLayer/ALIFLayer.forward(x (B,F), state) -> (Z, state) through the kernels with T = 1."""
	from snnimageclassification_b200 import ALIFLayer
	torch.manual_seed(3)
	lay = ALIFLayer(20, 32, use_recurrent_connection=True, device=DEV, learn_beta=False)
	x = (torch.rand(4, 5, 20, generator=torch.Generator().manual_seed(1)) < 0.4).float().to(DEV)
	state = None
	Zs, Vs = [], []
	for t in range(5):
		z, state = lay(x[:, t], state)
		assert len(state) == 3 and z is state[2]
		Zs.append(z); Vs.append(state[0])
	cfg = OracleCfg(4, 5, 20, 32, 1, 1, 0, 1, alpha=float(lay.alpha), rho=float(lay.rho), theta=float(lay.threshold),
		gamma=float(lay.gamma), kappa=0.0, beta=float(lay.beta))
	f = oracle.forward(cfg, npy(x), npy(lay.forward_weights), npy(lay.recurrent_weights), npy(lay.rec_mask),
		np.zeros((32, 1), np.float32), np.zeros(1, np.float32))
	assert np.array_equal(npy(torch.stack(Zs, 1)), f["Z"]) and np.array_equal(npy(torch.stack(Vs, 1)), f["V"])


def test_exec_batch_and_fit_on_synthetic_data(tmp_path):
	"""The reference's training entry points run end to end and the loss goes down."""
	from snnimageclassification_b200.datasets.datasets import SyntheticSpikeImages
	torch.manual_seed(0)
	T = 20
	enc = ToSpikes(T, use_periods=True, tau=20.0)
	net = SNN(784, 10, 128, use_recurrent_connection=True, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True, input_encoder=enc,
		checkpoint_folder=str(tmp_path / "ck"))
	ds = SyntheticSpikeImages(256, seed=1)
	ds.labels = (ds.images[:, :392].sum(1) > ds.images[:, 392:].sum(1)).long()   # a learnable rule
	train = torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)
	opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
	net.train()
	first = net._exec_batch(ds.images[:64], ds.labels[:64], torch.nn.NLLLoss(), opt)
	assert isinstance(first, float) and np.isfinite(first)
	hist = net.fit(train, train, nb_epochs=3, force_overwrite=True, verbose=False)
	assert len(hist["train"]) == 3 and hist["train"][-1] < first
	acc = net.compute_classification_accuracy(train)
	assert 0.0 <= acc <= 1.0
	# the validation pass of the last epoch counted the same accuracy (one pass instead of the reference's two)
	assert abs(net.last_eval_accuracy - acc) < 1e-12
	assert (tmp_path / "ck" / "snn-epoch2.pth").exists()
	# checkpoints written here load under torch's safe unpickler
	ck = torch.load(tmp_path / "ck" / "snn-epoch2.pth", map_location="cpu", weights_only=True)
	assert set(ck) >= {"epoch", "model_state_dict", "optimizer_state_dict"} or len(ck) >= 4
	# resume: nothing left to do, history restored from the checkpoints
	from snnimageclassification_b200 import LoadCheckpointMode
	hist2 = net.fit(train, train, nb_epochs=3, load_checkpoint_mode=LoadCheckpointMode.LAST_EPOCH, verbose=False)
	assert len(hist2["train"]) == 3


# ---- wide hidden layers (BASELINE configs[3], [4]) and non-power-of-two widths -------------------------------------
@pytest.mark.parametrize("H,B,T,N,layer,rec", [
	(256, 6, 10, 64, 1, 1), (512, 5, 9, 40, 0, 1), (1024, 5, 6, 784, 1, 1), (2048, 3, 5, 32, 1, 1), (384, 4, 7, 20, 1, 0)])
def test_wide_hidden_vs_oracle(H, B, T, N, layer, rec):
	"""H > 128 goes through recur_gen.cuh: same summation orders, so the fp32 forward is still bit-identical."""
	O = 10
	rng = np.random.default_rng(H + B)
	theta = 0.03 if layer else 1.0
	W_in = (rng.standard_normal((N, H)) * theta).astype(np.float32)
	W_rec = (rng.standard_normal((H, H)) * theta / 4).astype(np.float32) if rec else None
	mask = (1 - np.eye(H)).astype(np.float32) if rec else None
	W_out = (rng.standard_normal((H, O)) / 8).astype(np.float32)
	b_out = (rng.standard_normal(O) * 0.1).astype(np.float32)
	x = (rng.random((B, T, N)) < (0.1 if layer else 0.03)).astype(np.float32)
	labels = rng.integers(0, O, B)
	cfg = OracleCfg(B, T, N, H, O, layer_type=layer, surrogate=0, recurrent=rec, alpha=float(np.float32(np.exp(-1 / 20))),
		rho=float(np.float32(np.exp(-1 / 200))), theta=theta, gamma=0.3 if layer else 1.0,
		kappa=float(np.float32(np.exp(-1 / 10))), beta=1.6)
	out, beta = _gpu_forward(cfg, x, W_in, W_rec, mask, W_out, b_out)
	f = oracle.forward(cfg, x, W_in, W_rec, mask, W_out, b_out)
	for k in ("I_in", "V", "Z", "y") + (("a",) if layer else ()):
		assert np.array_equal(npy(out[k]), f[k]), k
	assert 0.0005 < f["Z"].mean() < 0.999
	h = oracle.head(f["y"], labels)
	assert np.array_equal(npy(out["logits"]), h["logits"]) and np.array_equal(npy(out["tstar"]), h["tstar"])
	loss, logp, g_logits = F_.run_head_nll(out["logits"], cu(labels, torch.int64))
	gref = oracle.backward(cfg, x, W_rec, mask, W_out, f["V"], f["a"], f["Z"], h["g_y"])
	for tc in (False, True):
		c = F_.LayerConsts(cfg.layer_type, cfg.surrogate, bool(cfg.recurrent), cfg.alpha, cfg.rho, cfg.theta, cfg.gamma,
			cfg.kappa, tensor_core=tc)
		g = F_.run_backward(c, cu(x), cu(W_rec), cu(mask), beta, cu(W_out), out["V"], out["a"], out["zbits"],
			g_logits=g_logits, tstar=out["tstar"], Z=out["Z"])
		assert rel_err(npy(g["gI"]()), gref["gI"]) <= 1e-5, tc
		for k in ("dW_in", "dW_out", "db") + (("dW_rec",) if rec else ()):
			assert rel_err(npy(g[k]), gref[k]) <= 1e-4, (tc, k)
	# tensor-core projection on the wide path
	c = F_.LayerConsts(cfg.layer_type, cfg.surrogate, bool(cfg.recurrent), cfg.alpha, cfg.rho, cfg.theta, cfg.gamma,
		cfg.kappa, tensor_core=True)
	o2 = F_.run_forward(c, cu(x), cu(W_in), cu(W_rec), cu(mask), beta, cu(W_out), cu(b_out))
	assert rel_err(npy(o2["I_in"]), f["I_in"]) <= 1e-5


@pytest.mark.parametrize("H", [100, 200, 20])
def test_hidden_widths_of_the_reference_sweeps(H):
	"""n_hidden_neurons in {100, 200} (training.py:37) are zero-padded to a supported width inside the glue."""
	torch.manual_seed(1)
	net = SNN(40, 10, H, use_recurrent_connection=True, int_time_steps=12, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=False, tensor_core=False)
	L, R = net.layers["input"], net.layers["readout"]
	assert tuple(L.forward_weights.shape) == (40, H) and tuple(L.recurrent_weights.shape) == (H, H)
	port = TorchPortSNN(40, H, 10, 12, layer_type=1, surrogate=0, recurrent=True, seed=0)
	port.load(npy(L.forward_weights), npy(L.recurrent_weights), npy(R.forward_weights), npy(R.bias_weights), beta=float(L.beta))
	g = torch.Generator().manual_seed(4)
	x = (torch.rand(5, 12, 40, generator=g) < 0.2).float()
	labels = torch.randint(0, 10, (5,), generator=g)
	net.train()
	y, hs = net(x)
	assert tuple(hs["input"][0].shape) == (5, 12, H) and tuple(hs["input"][2].shape) == (5, 12, H)
	p_out, p_hs = port.forward(x)
	assert np.array_equal(npy(hs["input"][2]), npy(p_hs["input"][2]))
	assert rel_err(npy(y), npy(p_out)) <= 1e-5
	loss = net.batch_loss(x, labels)
	net.zero_grad()
	loss.backward()
	p_loss = port.exec_batch(x, labels)
	assert abs(loss.item() - p_loss) <= 1e-5 * abs(p_loss)
	assert tuple(L.forward_weights.grad.shape) == (40, H)
	for mine, theirs in ((L.forward_weights, port.W_in), (L.recurrent_weights, port.W_rec), (R.forward_weights, port.W_out),
			(R.bias_weights, port.b_out)):
		assert rel_err(npy(mine.grad), npy(theirs.grad)) <= 1e-4


# ---- stacked hidden layers: the reference's own state_dict loaded into the B200 SNN ------------------------------------
STACKED = [str(n) for n in load("stacked_golden.npz")["names"]]


@pytest.mark.parametrize("name", STACKED)
@pytest.mark.parametrize("tc", [False, True])
def test_stacked_hidden_layers_vs_reference(name, tc):
	"""n_hidden_neurons=[h1, h2] (snn.py:116-128): load the reference's checkpoint tensors by name, run the same
	batch, compare every layer's raster, the output trace, the loss and every parameter gradient with the reference."""
	z = load("stacked_golden.npz")
	d = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
	alif, rec = (int(v) for v in d["flags"])
	widths = [int(v) for v in d["widths"]]
	kw = dict(learn_beta=False) if alif else {}
	net = SNN(48, 10, widths, use_recurrent_connection=bool(rec), int_time_steps=14, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=LayerType.ALIF if alif else LayerType.LIF, device=DEV, tensor_core=tc, **kw)
	assert list(net.state_dict().keys()) == [str(k) for k in d["keys"]]
	net.load_state_dict({str(k): torch.from_numpy(d[f"sd/{k}"]) for k in d["keys"]}, strict=True)
	x = torch.from_numpy(d["x"].astype(np.float32))
	labels = torch.from_numpy(d["labels"])
	net.train()
	logp, y, hs = net.get_prediction_log_proba(x, re_outputs_trace=True, re_hidden_states=True)
	assert list(hs.keys()) == ["input", "hidden_0", "readout"]
	for lname in ("input", "hidden_0"):
		assert np.array_equal(npy(hs[lname][-1]), d[f"Z/{lname}"]), lname          # rasters identical
		assert rel_err(npy(hs[lname][0]), d[f"V/{lname}"]) <= 1e-5, lname
	assert rel_err(npy(y), d["y"]) <= 1e-5
	loss = torch.nn.functional.nll_loss(logp, labels.to(DEV))
	assert abs(loss.item() - float(d["loss"])) <= 1e-5 * abs(float(d["loss"]))
	net.zero_grad()
	loss.backward()
	generic = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
	net.zero_grad()
	net.batch_loss(x, labels).backward()                                              # fused head on the last layer
	for k, p in net.named_parameters():
		assert p.grad is not None, k
		assert rel_err(npy(p.grad), d[f"grad/{k}"]) <= 1e-4, k
		assert rel_err(npy(generic[k]), d[f"grad/{k}"]) <= 1e-4, k
	with torch.no_grad():
		net.eval()
		lg = net.get_prediction_logits(x, re_outputs_trace=False, re_hidden_states=False)
	assert rel_err(npy(lg), d["y"].max(axis=1)) <= 1e-5


def test_fused_adam_matches_torch_adam():
	"""snnk_adam_step == torch.optim.Adam(lr, weight_decay=1e-5) (snn.py:299); parameters without gradient are skipped;
	the state_dict is interchangeable."""
	from snnimageclassification_b200 import FusedAdam
	g = torch.Generator().manual_seed(0)
	shapes = [(784, 128), (128, 128), (), (128, 10), (10,)]
	ref_p = [torch.randn(s, generator=g).to(DEV).requires_grad_() for s in shapes]
	my_p = [p.detach().clone().requires_grad_() for p in ref_p]
	ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=1e-5)
	mine = FusedAdam(my_p, lr=1e-3, weight_decay=1e-5)
	for it in range(6):
		for k, (a, b) in enumerate(zip(ref_p, my_p)):
			if k == 2:
				continue                      # the never-trained beta: grad stays None
			gr = torch.randn(a.shape, generator=g).to(DEV) * (10.0 if it == 3 else 1.0)
			a.grad, b.grad = gr.clone(), gr.clone()
		ref.step(); mine.step()
	for a, b in zip(ref_p, my_p):
		assert rel_err(npy(b), npy(a)) <= 1e-6
	assert torch.equal(ref_p[2], my_p[2])
	sd = mine.state_dict()
	assert float(sd["state"][0]["step"]) == 6.0 and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
	other = torch.optim.Adam([p.detach().clone().requires_grad_() for p in my_p], lr=1e-3, weight_decay=1e-5)
	other.load_state_dict(sd)                 # checkpoints interchange with the reference's optimizer


def test_fused_adam_loads_plain_adam_state_then_steps():
	"""The other direction of the interchange: a torch.optim.Adam / reference-checkpoint state -- non-capturable
	groups, `step` on the CPU under map_location='cpu' -- loaded into FusedAdam, followed by step()."""
	from snnimageclassification_b200 import FusedAdam
	g = torch.Generator().manual_seed(1)
	shapes = [(40, 32), (32, 32), (32, 10), (10,)]
	ref_p = [torch.randn(s, generator=g).to(DEV).requires_grad_() for s in shapes]
	ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=1e-5)
	grads = [[torch.randn(s, generator=g).to(DEV) for s in shapes] for _ in range(5)]
	for it in range(3):
		for p, gr in zip(ref_p, grads[it]):
			p.grad = gr.clone()
		ref.step()
	import io
	buf = io.BytesIO()
	torch.save(ref.state_dict(), buf)
	buf.seek(0)
	sd = torch.load(buf, map_location="cpu")            # what a resume under map_location='cpu' hands over
	assert sd["state"][0]["step"].device.type == "cpu" and not sd["param_groups"][0]["capturable"]
	my_p = [p.detach().clone().requires_grad_() for p in ref_p]
	mine = FusedAdam(my_p, lr=1e-3, weight_decay=1e-5)
	mine.load_state_dict(sd)
	assert all(gr["capturable"] and gr["foreach"] is False for gr in mine.param_groups)
	for p in my_p:
		st = mine.state[p]
		assert st["step"].device == p.device and st["step"].dtype == torch.float32 and st["step"].ndim == 0
		assert st["exp_avg"].device == p.device
	for it in range(3, 5):
		for a, b, gr in zip(ref_p, my_p, grads[it]):
			a.grad, b.grad = gr.clone(), gr.clone()
		ref.step(); mine.step()
	for a, b in zip(ref_p, my_p):
		assert rel_err(npy(b), npy(a)) <= 1e-6
	assert float(mine.state[my_p[0]]["step"]) == 5.0


def test_graphed_step_follows_lr_changes():
	"""lr / betas / eps / weight decay are kernel scalars frozen into a captured graph: the graphed step re-captures
	when they change, so a scheduler behaves as in eager mode."""
	from snnimageclassification_b200 import FusedAdam, LayerType, SNN, SpikeFuncType
	def make():
		torch.manual_seed(4)
		net = SNN(64, 10, 32, use_recurrent_connection=True, int_time_steps=12, spike_func=SpikeFuncType.FastSigmoid,
			hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True)
		net.train()
		return net, FusedAdam(net.parameters(), lr=1e-2, weight_decay=1e-5)
	g = torch.Generator().manual_seed(9)
	xs = [(torch.rand(16, 12, 64, generator=g) < 0.2).float() for _ in range(6)]
	ys = [torch.randint(0, 10, (16,), generator=g) for _ in range(6)]
	crit = torch.nn.NLLLoss()
	(eager, oe), (graphed, og) = make(), make()
	eager.cuda_graphs = False
	for it, (x, y) in enumerate(zip(xs, ys)):
		if it == 3:
			for o in (oe, og):
				o.param_groups[0]["lr"] = 1e-4
		eager._exec_batch(x, y, crit, oe)
		graphed._exec_batch(x, y, crit, og)
	for pe, pg in zip(eager.parameters(), graphed.parameters()):
		assert rel_err(npy(pg), npy(pe)) <= 1e-5


def test_head_label_semantics():
	"""ignore_index rows are excluded from the mean and get no gradient; other out-of-range labels poison the loss."""
	from snnimageclassification_b200.modules import functional as F_
	g = torch.Generator().manual_seed(2)
	logits = torch.randn(37, 10, generator=g).to(DEV)
	labels = torch.randint(0, 10, (37,), generator=g)
	labels[3] = labels[20] = -100
	loss, logp, gl = F_.run_head_nll(logits, labels.to(DEV))
	ref_lp = torch.log_softmax(logits.cpu(), -1).requires_grad_()
	lg = logits.cpu().clone().requires_grad_()
	ref = torch.nn.functional.nll_loss(torch.log_softmax(lg, -1), labels)
	ref.backward()
	assert abs(loss.item() - ref.item()) <= 1e-6 * abs(ref.item())
	assert rel_err(npy(gl), lg.grad.numpy()) <= 1e-6 and float(gl[3].abs().sum()) == 0.0
	labels[5] = 12
	loss, _, gl = F_.run_head_nll(logits, labels.to(DEV))
	assert torch.isnan(loss).item() and torch.isnan(gl[5]).all().item() and not torch.isnan(gl[4]).any().item()


# ---- fused head: snnk_forward_nll == snnk_forward + snnk_head_nll ------------------------------------------------------
@pytest.mark.parametrize("B,T,H,layer,rec,tc", [
	(256, 20, 128, 1, True, True),      # headline kernel family: register-resident recurrence, head in its tail
	(300, 7, 128, 1, True, False),      # more rows than the stand-alone kernel's 256 threads
	(1, 5, 32, 0, True, False), (37, 9, 64, 1, True, True),
	(1100, 3, 128, 0, True, False),     # two rows per CTA
	(64, 11, 128, 1, False, True),      # non-recurrent scan kernels: stand-alone head behind the forward kernels
	(40, 6, 256, 1, True, True),        # wide layer: stand-alone head
])
def test_fused_head_is_bit_identical_to_the_standalone_head(B, T, H, layer, rec, tc):
	N, O = 64, 10
	g = torch.Generator().manual_seed(B + H)
	theta = 0.03 if layer else 1.0
	x = F_.mark_binary((torch.rand(B, T, N, generator=g) < 0.3).float().to(DEV))
	W_in = (torch.randn(N, H, generator=g) * theta).to(DEV)
	W_rec = (torch.randn(H, H, generator=g) * theta).to(DEV) if rec else None
	mask = (1 - torch.eye(H)).to(DEV) if rec else None
	W_out, b_out = torch.randn(H, O, generator=g).to(DEV), (torch.randn(O, generator=g) * 0.1).to(DEV)
	beta = torch.tensor([1.6], device=DEV) if layer else None
	c = F_.LayerConsts(layer, 0, rec, 0.95, 0.995, theta, 0.3 if layer else 1.0, 0.9, tensor_core=tc)
	labels = torch.randint(0, O, (B,), generator=g).to(DEV)
	for variant in ("plain", "ignored", "bad"):
		lab = labels.clone()
		if variant == "ignored":
			lab[::3] = -100
		if variant == "bad" and B > 1:
			lab[B // 2] = O + 3
		for _ in range(2):      # twice: the ticket word must be back at zero after a launch
			f = F_.run_forward(c, x, W_in, W_rec, mask, beta, W_out, b_out, labels=lab)
		f0 = F_.run_forward(c, x, W_in, W_rec, mask, beta, W_out, b_out)
		loss, logp, gl = F_.run_head_nll(f0["logits"], lab)
		assert torch.equal(f["logits"], f0["logits"])
		assert torch.equal(f["logp"], logp)
		assert torch.equal(f["loss"].isnan(), loss.isnan()) and (bool(loss.isnan()) or torch.equal(f["loss"], loss)), variant
		assert torch.equal(f["g_logits"].isnan(), gl.isnan())
		assert torch.equal(torch.nan_to_num(f["g_logits"]), torch.nan_to_num(gl)), variant
		if variant == "bad" and B > 1:
			assert bool(f["loss"].isnan())


# ---- lean kernels of the headline geometry == the general kernels, bit for bit -----------------------------------------
@pytest.mark.parametrize("B,T,layer,phi,tc,dedup", [
	(256, 100, 1, 0, True, True),       # the bench workload's kernel variants: ALIF, tensor-core GEMMs, run table
	(256, 100, 1, 0, True, False),
	(300, 37, 0, 1, False, False),      # LIF, Phi, fp32 GEMMs, T not a multiple of the 8-step chunk
	(19, 1, 1, 1, True, False),         # T = 1
	(64, 128, 0, 0, True, True),        # longest sequence the lean BPTT sweep takes
])
def test_lean_kernels_match_general_kernels(B, T, layer, phi, tc, dedup, monkeypatch):
	"""recur_lean.cuh (default for recurrent LIF / ALIF, H = 128, one row per CTA) restates k_recur_fwd / k_recur_bwd with
	the bookkeeping taken out of the step loop: every output must be IDENTICAL (SNNK_LEAN=0 selects the general kernels)."""
	N, H, O = 784, 128, 10
	g = torch.Generator().manual_seed(B + T)
	theta = 0.03 if layer else 1.0
	if dedup:
		img = (torch.randint(1, 256, (B, N), generator=g).float() / 255.0) * (torch.rand(B, N, generator=g) < 0.19)
		x = ToSpikes(T, use_periods=True).encode_batch(img.to(DEV))
		assert F_.get_runs(x) is not None
	else:
		x = F_.mark_binary((torch.rand(B, T, N, generator=g) < (0.1 if layer else 0.02)).float().to(DEV))
	W_in = (torch.randn(N, H, generator=g) * theta).to(DEV)
	W_rec = (torch.randn(H, H, generator=g) * theta).to(DEV)
	mask = (1 - torch.eye(H)).to(DEV)
	W_out, b_out = torch.randn(H, O, generator=g).to(DEV), (torch.randn(O, generator=g) * 0.1).to(DEV)
	beta = torch.tensor([1.6], device=DEV) if layer else None
	labels = torch.randint(0, O, (B,), generator=g).to(DEV)
	c = F_.LayerConsts(layer, phi, True, 0.95, 0.995, theta, 0.3 if layer else 1.0, 0.9, tensor_core=tc)
	res = {}
	for lean in ("1", "0"):
		monkeypatch.setenv("SNNK_LEAN", lean)
		f = F_.run_forward(c, x, W_in, W_rec, mask, beta, W_out, b_out, labels=labels)
		gr = F_.run_backward(c, x, W_rec, mask, beta, W_out, f["V"], f["a"], f["zbits"], g_logits=f["g_logits"],
			tstar=f["tstar"], Z=f["Z"], W_effT=f["W_effT"])
		f2 = F_.run_forward(c, x, W_in, W_rec, mask, beta, W_out, b_out, traces=False)
		res[lean] = dict(V=f["V"], a=f["a"], Z=f["Z"], y=f["y"], zbits=f["zbits"], logits=f["logits"], tstar=f["tstar"],
			loss=f["loss"], logp=f["logp"], g_logits=f["g_logits"], gI=gr["gI"]().clone(), dW_in=gr["dW_in"], dW_rec=gr["dW_rec"],
			dW_out=gr["dW_out"], db=gr["db"], logits_infer=f2["logits"])
	assert float(res["1"]["Z"].mean()) > 0.001
	for k, v in res["1"].items():
		if v is not None:
			assert torch.equal(v, res["0"][k]), k
