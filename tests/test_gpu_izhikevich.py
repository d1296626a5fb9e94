"""IzhikevichLayer (the reference's third LayerType member, spiking_layers.py:246-353) on the GPU:
against the reference's own outputs (tests/golden/izhikevich_golden.npz) and bit-level against the C oracle."""
import numpy as np
import pytest
import torch

import oracle
from oracle import OracleCfg
from _util import dynamics_case, load, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
CASES = ["IZH_FastSigmoid_rec0", "IZH_FastSigmoid_rec1", "IZH_Phi_rec0", "IZH_Phi_rec1"]


def npy(t):
	return t.detach().cpu().numpy()


def _net(c, tensor_core):
	from snnimageclassification_b200 import LayerType, SNN, SpikeFuncType
	B, T, N, H, O = (int(v) for v in c["dims"])
	k = c["consts"]
	net = SNN(N, O, H, use_recurrent_connection=bool(c["flags"][1]), int_time_steps=T, dt=float(k[0]),
		spike_func=SpikeFuncType.Phi if c["flags"][0] else SpikeFuncType.FastSigmoid,
		hidden_layer_type=LayerType.Izhikevich, device=DEV, tensor_core=tensor_core)
	L, R = net.layers["input"], net.layers["readout"]
	with torch.no_grad():
		L.forward_weights.copy_(torch.from_numpy(c["W_in"]))
		if "W_rec" in c:
			L.recurrent_weights.copy_(torch.from_numpy(c["W_rec"]))
		R.forward_weights.copy_(torch.from_numpy(c["W_out"]))
		R.bias_weights.copy_(torch.from_numpy(c["b_out"]))
	for i, name in enumerate(("dt", "C", "v_rest", "v_th", "k", "a", "b", "c", "d", "v_peak", "gamma")):
		if name != "dt":
			assert abs(float(getattr(L, name)) - float(k[i])) < 1e-6, name    # the reference's defaults
	assert abs(float(R.kappa) - float(k[11])) < 1e-6
	return net


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tensor_core", [False, True])
def test_izhikevich_matches_reference(name, tensor_core):
	c = dynamics_case(load("izhikevich_golden.npz"), name)
	net = _net(c, tensor_core)
	x = torch.from_numpy(c["x"].astype(np.float32)).to(DEV)
	y = torch.from_numpy(c["labels"]).to(DEV)
	net.train()
	logp, out, hs = net.get_prediction_log_proba(x, re_outputs_trace=True, re_hidden_states=True)
	V, u, Z = hs["input"]
	assert np.array_equal(npy(Z).astype(np.uint8), c["Z"])
	# fp32 kernels: the north-star 1e-5.  Tensor-core GEMMs keep two tf32 planes of W_in (22 of 24 mantissa bits): with
	# this fixture's large weights (|w| up to ~100, currents of several hundred) that is 2e-5 of the membrane range.
	tol = 5e-5 if tensor_core else 1e-5
	assert rel_err(npy(V), c["V"]) <= tol and rel_err(npy(u), c["u"]) <= tol and rel_err(npy(out), c["y"]) <= tol
	assert rel_err(npy(logp), c["logp"]) <= tol
	net.zero_grad()
	loss = net.batch_loss(x, y, torch.nn.NLLLoss())
	loss.backward()
	assert abs(float(loss.detach()) - float(c["loss"])) <= tol * abs(float(c["loss"]))
	L, R = net.layers["input"], net.layers["readout"]
	assert rel_err(npy(L.forward_weights.grad), c["dW_in"]) <= 1e-4
	assert rel_err(npy(R.forward_weights.grad), c["dW_out"]) <= 1e-4 and rel_err(npy(R.bias_weights.grad), c["db"]) <= 1e-4
	if "dW_rec" in c:
		assert rel_err(npy(L.recurrent_weights.grad), c["dW_rec"]) <= 1e-4
	# the generic (any criterion) autograd path agrees with the fused-head one
	net.zero_grad()
	logp2, _, _ = net.get_prediction_log_proba(x, re_outputs_trace=True, re_hidden_states=True)
	torch.nn.functional.nll_loss(logp2, y).backward()
	assert rel_err(npy(L.forward_weights.grad), c["dW_in"]) <= 1e-4


@pytest.mark.parametrize("name", CASES)
def test_izhikevich_fp32_kernels_bit_identical_to_oracle(name):
	c = dynamics_case(load("izhikevich_golden.npz"), name)
	B, T, N, H, O = (int(v) for v in c["dims"])
	k = c["consts"]
	cfg = OracleCfg(B, T, N, H, O, layer_type=2, surrogate=int(c["flags"][0]), recurrent=int(c["flags"][1]),
		gamma=float(k[10]), kappa=float(k[11]), dt=float(k[0]), iz_C=float(k[1]), iz_vr=float(k[2]), iz_vth=float(k[3]),
		iz_k=float(k[4]), iz_a=float(k[5]), iz_b=float(k[6]), iz_c=float(k[7]), iz_d=float(k[8]), iz_vpeak=float(k[9]))
	xn = c["x"].astype(np.float32)
	f = oracle.forward(cfg, xn, c["W_in"], c.get("W_rec"), c.get("rec_mask"), c["W_out"], c["b_out"])
	net = _net(c, tensor_core=False)
	out, hs = net(torch.from_numpy(xn).to(DEV))
	V, u, Z = hs["input"]
	assert np.array_equal(npy(V), f["V"]) and np.array_equal(npy(u), f["a"]) and np.array_equal(npy(Z), f["Z"])
	assert np.array_equal(npy(out), f["y"])


def test_izhikevich_single_step_and_training():
	"""Layer-level forward(x, state) -> (Z, (V, u, Z)) as in the reference (:330-353), and a few training steps through
	_exec_batch (CUDA graph from the second batch on) reduce the loss."""
	from snnimageclassification_b200 import FusedAdam, LayerType, SNN, SpikeFuncType
	from snnimageclassification_b200.modules.spiking_layers import IzhikevichLayer
	torch.manual_seed(0)
	layer = IzhikevichLayer(20, 32, use_recurrent_connection=True, dt=1.0, device=DEV)
	with torch.no_grad():
		layer.forward_weights.mul_(30.0).add_(25.0)
	x = (torch.rand(4, 20, generator=torch.Generator().manual_seed(1)) < 0.5).float().to(DEV)
	state = None
	seen = 0.0
	for _ in range(30):
		z, state = layer(x, state)
		assert len(state) == 3 and z.shape == (4, 32)
		seen += float(z.sum())
	assert seen > 0
	torch.manual_seed(0)
	net = SNN(48, 10, 64, use_recurrent_connection=True, int_time_steps=30, dt=1.0, hidden_layer_type=LayerType.Izhikevich,
		spike_func=SpikeFuncType.FastSigmoid, device=DEV)
	with torch.no_grad():
		net.layers["input"].forward_weights.mul_(25.0).add_(20.0)
	g = torch.Generator().manual_seed(2)
	xb = (torch.rand(16, 30, 48, generator=g) < 0.2).float()
	yb = torch.randint(0, 10, (16,), generator=g)
	opt = FusedAdam(net.parameters(), lr=1e-2)
	net.train()
	losses = [net._exec_batch(xb, yb, torch.nn.NLLLoss(), opt) for _ in range(12)]
	assert all(np.isfinite(losses)) and losses[-1] < losses[0]


@pytest.mark.parametrize("H,rec,surr", [(256, True, 0), (200, True, 1), (384, False, 0)])
def test_izhikevich_wide_layers_vs_oracle(H, rec, surr):
	"""The reference puts no limit on the width of an IzhikevichLayer (spiking_layers.py:246-353); wider than 128 it runs
	on the fp32 kernels of recur_gen.cuh (200 is zero-padded to 256): traces bit-identical to the C oracle, gradients of
	the oracle's sweep within 1e-4."""
	from snnimageclassification_b200 import LayerType, SNN, SpikeFuncType
	B, T, N, O = 6, 25, 48, 10
	torch.manual_seed(3)
	net = SNN(N, O, H, use_recurrent_connection=rec, int_time_steps=T, dt=1.0, hidden_layer_type=LayerType.Izhikevich,
		spike_func=SpikeFuncType.Phi if surr else SpikeFuncType.FastSigmoid, device=DEV, tensor_core=False)
	L, R = net.layers["input"], net.layers["readout"]
	with torch.no_grad():
		L.forward_weights.mul_(25.0).add_(20.0)
		if rec:
			L.recurrent_weights.mul_(5.0)
	g = torch.Generator().manual_seed(4)
	x = (torch.rand(B, T, N, generator=g) < 0.2).float()
	lab = torch.randint(0, O, (B,), generator=g)
	# the oracle (like the kernels) runs the padded width: zero columns / rows for the neurons that do not exist
	Hp = (H + 127) // 128 * 128
	cfg = OracleCfg(B, T, N, Hp, O, layer_type=2, surrogate=surr, recurrent=int(rec), gamma=float(L.gamma), kappa=float(R.kappa),
		dt=float(L.dt), iz_C=float(L.C), iz_vr=float(L.v_rest), iz_vth=float(L.v_th), iz_k=float(L.k), iz_a=float(L.a),
		iz_b=float(L.b), iz_c=float(L.c), iz_d=float(L.d), iz_vpeak=float(L.v_peak))

	def pad(a, rows, cols):
		out_ = np.zeros((rows, cols), np.float32)
		out_[:a.shape[0], :a.shape[1]] = a
		return out_
	W_in = pad(npy(L.forward_weights), N, Hp)
	W_rec = pad(npy(L.recurrent_weights), Hp, Hp) if rec else np.zeros((Hp, Hp), np.float32)
	mask = pad(npy(L.rec_mask), Hp, Hp) if rec else np.zeros((Hp, Hp), np.float32)
	W_out = pad(npy(R.forward_weights), Hp, O)
	f = oracle.forward(cfg, x.numpy(), W_in, W_rec if rec else None, mask if rec else None, W_out, npy(R.bias_weights))
	net.train()
	out, hs = net(x.to(DEV))
	V, u, Z = hs["input"]
	assert npy(Z).sum() > 0 and V.shape == (B, T, H)
	assert np.array_equal(npy(Z), f["Z"][..., :H]) and np.array_equal(npy(V), f["V"][..., :H])
	assert np.array_equal(npy(u), f["a"][..., :H]) and np.array_equal(npy(out), f["y"])
	net.zero_grad()
	loss = net.batch_loss(x.to(DEV), lab.to(DEV), torch.nn.NLLLoss())
	loss.backward()
	h = oracle.head(f["y"], lab.numpy())
	assert abs(loss.item() - h["loss"]) <= 1e-5 * abs(h["loss"])
	gr = oracle.backward(cfg, x.numpy(), W_rec, mask, W_out, f["V"], f["a"], f["Z"], h["g_y"])
	assert rel_err(npy(L.forward_weights.grad), gr["dW_in"][:, :H]) <= 1e-4
	assert rel_err(npy(R.forward_weights.grad), gr["dW_out"][:H]) <= 1e-4 and rel_err(npy(R.bias_weights.grad), gr["db"]) <= 1e-4
	if rec:
		assert rel_err(npy(L.recurrent_weights.grad), gr["dW_rec"][:H, :H]) <= 1e-4
