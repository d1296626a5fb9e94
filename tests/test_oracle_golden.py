"""Pins the CPU oracle (oracle/snn_oracle.c and oracle/torch_port.py) against the reference.

Golden data comes from tests/golden/make_golden.py, which ran the reference itself; the known-answer
arrays below are the ones the reference's own tests use (test/test_to_spikes.py, cited per test).
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import OracleCfg
from oracle.torch_port import TorchPortSNN
from _util import dynamics_case, first_divergence, load, rel_err, unpack_bits

KW = dict(tau=20.0, thr=0.2, eps=1e-7)


# ---- encoder: the reference's own known-answer tests -------------------------------------------------------------
def test_periods_zero_pixel():
	# test/test_to_spikes.py:9-13
	assert np.all(oracle.periods(np.array([0.0]), 100, 20.0, 0.2, 1e-7) == 100)


def test_periods_known_answer_1():
	# test/test_to_spikes.py:15-20
	pix = np.array([0.82352941, 0.82745098, 0.83529412, 0.8745098, 0.8627451, 0.95294118, 0.79215686, 0., 0., 0.])
	assert np.array_equal(oracle.periods(pix, 100, 20.0, 0.2, 1e-7), [5, 5, 5, 5, 5, 4, 5, 100, 100, 100])


def test_periods_known_answer_2():
	# test/test_to_spikes.py:22-30
	pix = np.array([0.8627451, 0.90980392, 0.96470588, 0., 0.01176471, 0.79215686, 0.89411765, 0.87843137,
		0.86666667, 0.82745098])
	assert np.array_equal(oracle.periods(pix, 10, 20.0, 0.2, 1e-7), [5, 4, 4, 10, 10, 5, 5, 5, 5, 5])


def _expected_call():
	exp = np.zeros((10, 12), dtype=np.uint8)
	exp[[4, 4, 5, 5, 5, 5, 5, 5, 5, 5], [1, 2, 0, 5, 6, 7, 8, 9, 10, 11]] = 1
	return exp


def test_call_known_answer():
	# test/test_to_spikes.py:38-50
	pix = np.array([0.8627451, 0.90980392, 0.96470588, 0., 0.01176471, 0.79215686, 0.89411765, 0.87843137,
		0.86666667, 0.82745098, 0.82745098, 0.83921569])
	assert np.array_equal(oracle.encode(pix, 10, 10, periodic=False, **KW), _expected_call())


def test_firing_times_to_spikes_known_answer():
	# test/test_to_spikes.py:52-60
	ft = np.array([[5, 4, 4, 10, 10, 5, 5, 5, 5, 5, 5, 5]])
	assert np.array_equal(oracle.raster(ft, 10, periodic=False)[0], _expected_call())


def test_firing_periods_to_spikes_known_answer():
	# test/test_to_spikes.py:62-73
	exp = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 0, 0], [1, 1, 1]], dtype=np.uint8)
	assert np.array_equal(oracle.raster(np.array([[1, 2, 6]]), 5, periodic=True)[0], exp)


def test_golden_image():
	# test/test_to_spikes.py:75-83 (golden vector regenerated from the reference into encoder_golden.npz)
	z = load("encoder_golden.npz")
	x = z["real_x_f64"]
	spikes = unpack_bits(z["real_spikes_bits"], tuple(z["real_spikes_shape"]))
	assert spikes.sum() == 390
	assert np.array_equal(oracle.periods(x, 100, 20.0, 0.2, 1e-7), z["real_periods"])
	assert np.array_equal(oracle.encode(x, 100, 100, periodic=False, **KW), spikes)


def test_encoder_against_reference_outputs():
	z = load("encoder_golden.npz")
	for key in z["cases"]:
		x = z[f"{key}_x"]
		parts = key.split("_")
		tau = float(parts[2][3:]); periodic = bool(int(parts[3][1:])); n = int(parts[4][1:])
		per = oracle.periods(x, n, tau, 0.2, 1e-7)
		assert np.array_equal(per, z[f"{key}_periods"]), key
		ras = oracle.encode(x, n, n, tau=tau, thr=0.2, periodic=periodic, eps=1e-7)
		assert np.array_equal(ras, unpack_bits(z[f"{key}_bits"], ras.shape)), key


# ---- dynamics, loss, BPTT ------------------------------------------------------------------------------------------
def _cfg(d):
	B, T, N, H, O = (int(v) for v in d["dims"])
	alif, phi, rec, _ = (int(v) for v in d["flags"])
	al, rho, th, ga, ka, be = (float(v) for v in d["scalars"])
	return OracleCfg(B, T, N, H, O, layer_type=alif, surrogate=phi, recurrent=rec, alpha=al, rho=rho, theta=th,
		gamma=ga, kappa=ka, beta=be)


NAMES = [str(n) for n in load("dynamics_golden.npz")["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_matches_reference(name):
	d = dynamics_case(load("dynamics_golden.npz"), name)
	cfg = _cfg(d)
	x = d["x"].astype(np.float32)
	f = oracle.forward(cfg, x, d["W_in"], d.get("W_rec"), d.get("rec_mask"), d["W_out"], d["b_out"])
	Zref = d["Z"].astype(np.float32)
	# spike rasters: the bar is >= 99.99 % identical; on these fixtures they are identical
	assert (f["Z"] == Zref).mean() >= 0.9999, name
	fd = first_divergence(f["Z"], Zref)
	for b in range(cfg.B):  # state parity (1e-5 relative) on the prefix where the rasters agree
		t = fd[b]
		scale = max(np.abs(d["V"][b, :t]).max(), 1e-6) if t else 1.0
		assert np.abs(f["V"][b, :t] - d["V"][b, :t]).max() <= 1e-5 * scale, name
		if cfg.layer_type == 1 and t:
			assert np.abs(f["a"][b, :t] - d["a"][b, :t]).max() <= 1e-5 * max(np.abs(d["a"][b, :t]).max(), 1e-6)
	if (fd == cfg.T).all():
		assert rel_err(f["y"], d["y"]) <= 1e-5
		h = oracle.head(f["y"], d["labels"])
		assert rel_err(h["logp"], d["logp"]) <= 1e-5
		assert abs(h["loss"] - float(d["loss"])) <= 1e-5 * abs(float(d["loss"]))
		g = oracle.backward(cfg, x, d.get("W_rec"), d.get("rec_mask"), d["W_out"], f["V"], f["a"], f["Z"], h["g_y"])
		assert rel_err(g["dW_in"], d["dW_in"]) <= 1e-4
		assert rel_err(g["dW_out"], d["dW_out"]) <= 1e-4
		assert rel_err(g["db"], d["db"]) <= 1e-4
		if cfg.recurrent:
			assert rel_err(g["dW_rec"], d["dW_rec"]) <= 1e-4
			assert np.all(np.diag(g["dW_rec"]) == 0.0)  # rec_mask zeroes the diagonal (spiking_layers.py:52)
	else:
		pytest.fail(f"{name}: oracle raster diverged from the reference fixture at steps {fd}")


@pytest.mark.parametrize("name", NAMES)
def test_torch_port_matches_reference(name):
	d = dynamics_case(load("dynamics_golden.npz"), name)
	cfg = _cfg(d)
	alif, phi, rec, lb = (int(v) for v in d["flags"])
	net = TorchPortSNN(cfg.N, cfg.H, cfg.O, cfg.T, layer_type=alif, surrogate=phi, recurrent=bool(rec),
		learn_beta=bool(lb), seed=0)
	net.load(d["W_in"], d.get("W_rec"), d["W_out"], d["b_out"], beta=cfg.beta)
	assert float(net.alpha) == pytest.approx(cfg.alpha, rel=1e-7)
	assert float(net.kappa) == pytest.approx(cfg.kappa, rel=1e-7)
	if alif:
		assert float(net.rho) == pytest.approx(cfg.rho, rel=1e-7)
	x = torch.from_numpy(d["x"].astype(np.float32))
	loss = net.exec_batch(x, torch.from_numpy(d["labels"]))
	assert abs(loss - float(d["loss"])) <= 1e-6 * abs(float(d["loss"]))
	assert rel_err(net.W_in.grad.numpy(), d["dW_in"]) <= 1e-5
	assert rel_err(net.W_out.grad.numpy(), d["dW_out"]) <= 1e-5
	assert rel_err(net.b_out.grad.numpy(), d["db"]) <= 1e-5
	if rec:
		assert rel_err(net.W_rec.grad.numpy(), d["dW_rec"]) <= 1e-5
	if alif and lb:
		# reference quirk: the threshold input of the spike function gets no gradient, so beta.grad is None
		assert int(d["beta_grad_is_none"]) == 1
		assert net.beta.grad is None


# ---- IzhikevichLayer (third LayerType member, spiking_layers.py:246-353) ---------------------------------------------
def _izh_cfg(c):
	from oracle import OracleCfg
	B, T, N, H, O = (int(v) for v in c["dims"])
	k = c["consts"]
	return OracleCfg(B, T, N, H, O, layer_type=2, surrogate=int(c["flags"][0]), recurrent=int(c["flags"][1]),
		gamma=float(k[10]), kappa=float(k[11]), dt=float(k[0]), iz_C=float(k[1]), iz_vr=float(k[2]), iz_vth=float(k[3]),
		iz_k=float(k[4]), iz_a=float(k[5]), iz_b=float(k[6]), iz_c=float(k[7]), iz_d=float(k[8]), iz_vpeak=float(k[9]))


@pytest.mark.parametrize("name", ["IZH_FastSigmoid_rec0", "IZH_FastSigmoid_rec1", "IZH_Phi_rec0", "IZH_Phi_rec1"])
def test_oracle_izhikevich_matches_reference(name):
	import oracle
	z = load("izhikevich_golden.npz")
	c = dynamics_case(z, name)
	cfg = _izh_cfg(c)
	x = c["x"].astype(np.float32)
	f = oracle.forward(cfg, x, c["W_in"], c.get("W_rec"), c.get("rec_mask"), c["W_out"], c["b_out"])
	assert float(c["spike_rate"]) > 0.01
	assert np.array_equal(f["Z"].astype(np.uint8), c["Z"])
	assert rel_err(f["V"], c["V"]) <= 1e-5 and rel_err(f["a"], c["u"]) <= 1e-5 and rel_err(f["y"], c["y"]) <= 1e-5
	h = oracle.head(f["y"], c["labels"])
	assert rel_err(h["logp"], c["logp"]) <= 1e-5 and abs(h["loss"] - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
	g = oracle.backward(cfg, x, c.get("W_rec"), c.get("rec_mask"), c["W_out"], f["V"], f["a"], f["Z"], h["g_y"])
	assert rel_err(g["dW_in"], c["dW_in"]) <= 1e-4
	assert rel_err(g["dW_out"], c["dW_out"]) <= 1e-4 and rel_err(g["db"], c["db"]) <= 1e-4
	if cfg.recurrent:
		assert rel_err(g["dW_rec"], c["dW_rec"]) <= 1e-4


@pytest.mark.parametrize("name", ["IZH_FastSigmoid_rec1", "IZH_Phi_rec0"])
def test_torch_port_izhikevich_matches_reference(name):
	z = load("izhikevich_golden.npz")
	c = dynamics_case(z, name)
	B, T, N, H, O = (int(v) for v in c["dims"])
	k = c["consts"]
	net = TorchPortSNN(N, H, O, T, layer_type=2, surrogate=int(c["flags"][0]), recurrent=bool(c["flags"][1]), dt=float(k[0]))
	net.load(c["W_in"], c.get("W_rec"), c["W_out"], c["b_out"])
	x = torch.from_numpy(c["x"].astype(np.float32))
	logp, out, hs = net.log_proba(x)
	V, u, Z = hs["input"]
	assert np.array_equal(Z.detach().numpy().astype(np.uint8), c["Z"])
	assert rel_err(V.detach().numpy(), c["V"]) <= 1e-5 and rel_err(u.detach().numpy(), c["u"]) <= 1e-5
	loss = torch.nn.functional.nll_loss(logp, torch.from_numpy(c["labels"]))
	loss.backward()
	assert abs(float(loss) - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
	assert rel_err(net.W_in.grad.numpy(), c["dW_in"]) <= 1e-4 and rel_err(net.W_out.grad.numpy(), c["dW_out"]) <= 1e-4
	if net.recurrent:
		assert rel_err(net.W_rec.grad.numpy(), c["dW_rec"]) <= 1e-4


# ---- property checks of the encoder oracle against a second, numpy restatement of datasets.py:42-86 ------------------
def _np_periods(x, t_max, tau, thr, eps):
	x = np.asarray(x)
	below = x < thr
	xc = np.clip(x, thr + eps, 1e9)
	T = tau * np.log(xc / (xc - thr))
	T[below] = t_max
	return T.astype(np.int64)


def _np_periodic(per, n_steps):
	p = np.clip(per, 1, n_steps - 1)
	t = np.arange(n_steps)[:, None]
	return ((t >= p[None, :]) & ((t - p[None, :]) % p[None, :] == 0)).astype(np.uint8)


def _np_latency(per, n_steps):
	t = np.arange(n_steps)[:, None]
	return (t == per[None, :]).astype(np.uint8)


def test_encoder_oracle_properties_float64():
	from hypothesis import given, settings, strategies as st

	@settings(max_examples=60, deadline=None)
	@given(st.integers(0, 2**31 - 1), st.sampled_from([2, 10, 32, 100]), st.sampled_from([20.0, 0.02, 1.0]))
	def check(seed, n_steps, tau):
		rng = np.random.default_rng(seed)
		x = rng.random(97)                      # float64: numpy's log is what the reference itself runs
		x[rng.random(97) < 0.3] = 0.0
		per = oracle.periods(x, n_steps, tau, 0.2, 1e-7)
		assert np.array_equal(per, _np_periods(x, n_steps, tau, 0.2, 1e-7))
		assert np.array_equal(oracle.raster(per[None], n_steps, True)[0], _np_periodic(per, n_steps))
		assert np.array_equal(oracle.raster(per[None], n_steps, False)[0], _np_latency(per, n_steps))
		r = oracle.raster(per[None], n_steps, True)[0].astype(np.int64)
		p = np.clip(per, 1, n_steps - 1)
		assert np.array_equal(r.sum(0), (n_steps - 1) // p)          # a pixel of period p fires floor((T-1)/p) times
		assert r[0].sum() == 0                                        # never at t = 0 (p >= 1)
	check()


def test_dynamics_oracle_structural_properties():
	"""Size-independent properties of the restated dynamics: batch rows are independent; zero input keeps a LIF layer
	silent; the head's gradient w.r.t. the output trace is non-zero only at the (first) arg-max step of each class."""
	rng = np.random.default_rng(5)
	B, T, N, H, O = 6, 20, 24, 32, 10
	cfg = OracleCfg(B, T, N, H, O, layer_type=1, surrogate=0, recurrent=1, alpha=0.95, rho=0.99, theta=0.03, gamma=0.3,
		kappa=0.9, beta=0.01)
	x = (rng.random((B, T, N)) < 0.2).astype(np.float32)
	W_in = (rng.standard_normal((N, H)) * 0.03).astype(np.float32)
	W_rec = (rng.standard_normal((H, H)) * 0.03).astype(np.float32)
	mask = (1 - np.eye(H)).astype(np.float32)
	W_out = rng.standard_normal((H, O)).astype(np.float32)
	b = np.zeros(O, np.float32)
	f = oracle.forward(cfg, x, W_in, W_rec, mask, W_out, b)
	perm = rng.permutation(B)
	fp = oracle.forward(cfg, x[perm], W_in, W_rec, mask, W_out, b)
	for k in ("V", "a", "Z", "y"):
		assert np.array_equal(fp[k], f[k][perm])
	lif = OracleCfg(B, T, N, H, O, layer_type=0, surrogate=0, recurrent=1, alpha=0.9, theta=1.0, gamma=1.0, kappa=0.9)
	z = oracle.forward(lif, np.zeros_like(x), W_in, W_rec, mask, W_out, b)
	assert not z["Z"].any() and not z["V"].any() and not z["y"].any()
	labels = rng.integers(0, O, B)
	h = oracle.head(f["y"], labels)
	assert np.allclose(np.exp(h["logp"]).sum(1), 1.0, atol=1e-6)
	nz = h["g_y"] != 0
	assert nz.sum(axis=1).max() <= 1                                  # at most one step per (sample, class)
	bi, ti, ci = np.nonzero(nz)
	assert np.array_equal(ti, h["tstar"][bi, ci])
	assert np.array_equal(h["logits"], f["y"].max(axis=1))
	assert np.array_equal(h["tstar"], f["y"].argmax(axis=1))           # first maximum wins (snn.py:228)
