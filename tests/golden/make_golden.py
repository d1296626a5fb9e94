"""Regenerates tests/golden/*.npz by RUNNING THE REFERENCE (read-only checkout at /root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
  encoder_golden.npz   ToSpikes outputs of the reference: its golden image (test/test_x_to_spikes.npy), the
                       256 pixel levels k/255, seeded random inputs; float32 and float64; tau in {20, 0.02};
                       periodic and non-periodic.  Rasters are stored bit-packed.
  dynamics_golden.npz  For every (LIF|ALIF) x (rec|non-rec) x (FastSigmoid|Phi) [x learn_beta]: inputs, the
                       reference's initial weights, forward traces, log-probabilities, loss and autograd
                       gradients from the reference's own SNN class.
  init_golden.npz      state_dict of reference SNNs built under torch.manual_seed (RNG-order parity).
  stacked_golden.npz   Two stacked hidden layers: the reference's state_dict, traces, loss and gradients.
  izhikevich_golden.npz  IzhikevichLayer x (rec|non-rec) x (FastSigmoid|Phi): traces (V, u, Z), loss, gradients.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SNN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from src.datasets.datasets import ToSpikes  # noqa: E402
from src.modules.snn import SNN  # noqa: E402
from src.modules.spike_funcs import SpikeFuncType  # noqa: E402
from src.modules.spiking_layers import LayerType  # noqa: E402


def pack(r):
	return np.packbits(np.asarray(r) != 0, axis=None)


def encoder_fixture():
	out = {}
	# (1) the reference's golden image: test/test_to_spikes.py:75-83
	d = np.load(os.path.join(REF, "test", "test_x_to_spikes.npy"), allow_pickle=True).item()
	x_img = np.asarray(d["x"], dtype=np.float64)  # (28, 28) in 0..255
	x = torch.flatten(torch.as_tensor(x_img)[None] / 255.0).numpy()  # float64, as ToTensor on a float64 HxW array
	ts = ToSpikes(100, 100, tau=20.0, thr=0.2, epsilon=1e-7)
	spikes = ts(x.copy()).numpy()
	assert np.allclose(spikes, d["spikes"]), "reference no longer reproduces its own golden vector"
	out["real_x_f64"] = x
	out["real_spikes_shape"] = np.array(spikes.shape)
	out["real_spikes_bits"] = pack(spikes)
	out["real_periods"] = ts.pixels_to_firing_periods(x.copy())

	# (2) pixel levels + random inputs, both dtypes, both tau regimes, both modes
	rng = np.random.default_rng(1234)
	levels = np.arange(256) / 255.0
	rand = rng.uniform(0.0, 1.0, size=(4, 97))
	rand[0, :8] = [0.0, 0.2, 0.2000001, 0.19999999, 1.0, 0.5, 0.21, 0.999]
	cases = []
	for name, arr in (("levels", levels[None]), ("rand", rand)):
		for dt in (np.float32, np.float64):
			for tau in (20.0, 0.02):
				for periodic in (False, True):
					for n_steps in (100, 10):
						key = f"{name}_{np.dtype(dt).name}_tau{tau}_p{int(periodic)}_n{n_steps}"
						xin = arr.astype(dt)
						ts = ToSpikes(n_steps, n_steps, tau=tau, thr=0.2, use_periods=periodic, epsilon=1e-7)
						per = np.stack([ts.pixels_to_firing_periods(r.copy()) for r in xin])
						ras = np.stack([ToSpikes(n_steps, n_steps, tau=tau, thr=0.2, use_periods=periodic, epsilon=1e-7)(
							r.copy()).numpy() for r in xin])
						assert ras.dtype == np.float64 and ras.shape == (xin.shape[0], n_steps, xin.shape[1])
						out[key + "_x"] = xin
						out[key + "_periods"] = per.astype(np.int64)
						out[key + "_bits"] = pack(ras)
						cases.append(key)
	out["cases"] = np.array(cases)
	np.savez_compressed(os.path.join(HERE, "encoder_golden.npz"), **out)
	print("encoder fixture:", len(cases), "cases")


def run_reference(layer, sf, rec, learn_beta, B, T, N, H, O, seed, density):
	torch.manual_seed(seed)
	kw = dict(learn_beta=learn_beta) if layer == LayerType.ALIF else {}
	net = SNN(
		N, O, H, use_recurrent_connection=rec, int_time_steps=T, spike_func=sf, hidden_layer_type=layer,
		device=torch.device("cpu"), **kw)
	g = torch.Generator().manual_seed(seed + 1)
	x = (torch.rand(B, T, N, generator=g) < density).float()
	labels = torch.randint(0, O, (B,), generator=g)
	net.train()
	logp, out, hs = net.get_prediction_log_proba(x, re_outputs_trace=True, re_hidden_states=True)
	loss = torch.nn.NLLLoss()(logp, labels)
	net.zero_grad()
	loss.backward()
	L = net.layers["input"]
	R = net.layers["readout"]
	d = dict(
		x=x.numpy().astype(np.uint8), labels=labels.numpy(), W_in=L.forward_weights.detach().numpy(),
		W_out=R.forward_weights.detach().numpy(), b_out=R.bias_weights.detach().numpy(),
		V=hs["input"][0].detach().numpy(), Z=hs["input"][-1].detach().numpy().astype(np.uint8),
		y=out.detach().numpy(), logp=logp.detach().numpy(), loss=np.float32(loss.item()),
		dW_in=L.forward_weights.grad.numpy(), dW_out=R.forward_weights.grad.numpy(), db=R.bias_weights.grad.numpy(),
		scalars=np.array([float(L.alpha), float(getattr(L, "rho", torch.tensor(0.0))), float(L.threshold),
			float(L.gamma), float(R.kappa), float(getattr(L, "beta", torch.tensor(0.0)))], dtype=np.float32),
		dims=np.array([B, T, N, H, O]),
		flags=np.array([int(layer == LayerType.ALIF), int(sf == SpikeFuncType.Phi), int(rec), int(learn_beta)]),
	)
	if layer == LayerType.ALIF:
		d["a"] = hs["input"][1].detach().numpy()
		beta = L.beta
		d["beta_grad_is_none"] = np.array(int(getattr(beta, "grad", None) is None))
	if rec:
		d["W_rec"] = L.recurrent_weights.detach().numpy()
		d["rec_mask"] = L.rec_mask.numpy()
		d["dW_rec"] = L.recurrent_weights.grad.numpy()
	return d


def dynamics_fixture():
	out = {}
	names = []
	i = 0
	for layer in (LayerType.LIF, LayerType.ALIF):
		for sf in (SpikeFuncType.FastSigmoid, SpikeFuncType.Phi):
			for rec in (False, True):
				for lb in ((False, True) if layer == LayerType.ALIF else (False,)):
					name = f"{layer.name}_{sf.name}_rec{int(rec)}_lb{int(lb)}"
					# LIF weights ~N(0,1) with theta=1: keep inputs sparse so the layer is not saturated
					d = run_reference(layer, sf, rec, lb, B=3, T=16, N=48, H=32, O=10, seed=100 + i,
						density=0.08 if layer == LayerType.LIF else 0.15)
					for k, v in d.items():
						out[f"{name}/{k}"] = v
					names.append(name)
					i += 1
	# one case at the headline geometry (784-128-10, T=100), small batch
	d = run_reference(LayerType.ALIF, SpikeFuncType.FastSigmoid, True, True, B=2, T=100, N=784, H=128, O=10,
		seed=7, density=0.1)
	for k, v in d.items():
		if k in ("dW_in", "W_in"):
			v = v.astype(np.float32)
		out[f"headline/{k}"] = v
	names.append("headline")
	out["names"] = np.array(names)
	np.savez_compressed(os.path.join(HERE, "dynamics_golden.npz"), **out)
	print("dynamics fixture:", names)


def init_fixture():
	out = {}
	specs = [
		("alif_rec_lb", dict(hidden_layer_type=LayerType.ALIF, use_recurrent_connection=True, learn_beta=True)),
		("alif_nonrec", dict(hidden_layer_type=LayerType.ALIF, use_recurrent_connection=False, learn_beta=False)),
		("lif_rec", dict(hidden_layer_type=LayerType.LIF, use_recurrent_connection=True)),
		("lif_two_hidden", dict(hidden_layer_type=LayerType.LIF, use_recurrent_connection=True)),
	]
	for name, kw in specs:
		torch.manual_seed(42)
		hidden = [24, 16] if name == "lif_two_hidden" else 24
		net = SNN(20, 10, hidden, int_time_steps=5, spike_func=SpikeFuncType.FastSigmoid,
			device=torch.device("cpu"), **kw)
		keys = []
		for k, v in net.state_dict().items():
			out[f"{name}/{k}"] = v.numpy()
			keys.append(k)
		out[f"{name}/__keys__"] = np.array(keys)
		out[f"{name}/__params__"] = np.array([n for n, _ in net.named_parameters()])
	out["names"] = np.array([s[0] for s in specs])
	np.savez_compressed(os.path.join(HERE, "init_golden.npz"), **out)
	print("init fixture done")


def izhikevich_fixture():
	"""IzhikevichLayer (spiking_layers.py:246-353), the third LayerType member.  With the reference's default constants
	(C=100, k=0.7, dt=1e-3) and N(0,1) weights the membrane barely leaves v_rest, so the cases use dt = 1.0 (the
	millisecond units the constants come from) and scaled-up weights: every neuron spikes a few times in T steps."""
	out = {}
	names = []
	i = 0
	for sf in (SpikeFuncType.FastSigmoid, SpikeFuncType.Phi):
		for rec in (False, True):
			name = f"IZH_{sf.name}_rec{int(rec)}"
			B, T, N, H, O = 3, 40, 48, 32, 10
			torch.manual_seed(300 + i)
			net = SNN(N, O, H, use_recurrent_connection=rec, int_time_steps=T, spike_func=sf,
				hidden_layer_type=LayerType.Izhikevich, dt=1.0, device=torch.device("cpu"))
			L, R = net.layers["input"], net.layers["readout"]
			with torch.no_grad():
				L.forward_weights.mul_(25.0).add_(20.0)
				if rec:
					L.recurrent_weights.mul_(10.0)
			g = torch.Generator().manual_seed(301 + i)
			x = (torch.rand(B, T, N, generator=g) < 0.2).float()
			labels = torch.randint(0, O, (B,), generator=g)
			net.train()
			logp, y, hs = net.get_prediction_log_proba(x, re_outputs_trace=True, re_hidden_states=True)
			loss = torch.nn.NLLLoss()(logp, labels)
			net.zero_grad()
			loss.backward()
			V, u, Z = hs["input"]
			d = dict(
				x=x.numpy().astype(np.uint8), labels=labels.numpy(), W_in=L.forward_weights.detach().numpy(),
				W_out=R.forward_weights.detach().numpy(), b_out=R.bias_weights.detach().numpy(),
				V=V.detach().numpy(), u=u.detach().numpy(), Z=Z.detach().numpy().astype(np.uint8), y=y.detach().numpy(),
				logp=logp.detach().numpy(), loss=np.float32(loss.item()),
				dW_in=L.forward_weights.grad.numpy(), dW_out=R.forward_weights.grad.numpy(), db=R.bias_weights.grad.numpy(),
				consts=np.array([float(L.dt), float(L.C), float(L.v_rest), float(L.v_th), float(L.k), float(L.a), float(L.b),
					float(L.c), float(L.d), float(L.v_peak), float(L.gamma), float(R.kappa)], dtype=np.float32),
				dims=np.array([B, T, N, H, O]), flags=np.array([int(sf == SpikeFuncType.Phi), int(rec)]),
				spike_rate=np.float32(Z.detach().mean().item()),
			)
			if rec:
				d["W_rec"] = L.recurrent_weights.detach().numpy()
				d["rec_mask"] = L.rec_mask.numpy()
				d["dW_rec"] = L.recurrent_weights.grad.numpy()
			for k, v in d.items():
				out[f"{name}/{k}"] = v
			names.append(name)
			print(name, "spike rate", float(d["spike_rate"]), "loss", float(d["loss"]), "|dW_in|max", float(np.abs(d["dW_in"]).max()))
			i += 1
	out["names"] = np.array(names)
	np.savez_compressed(os.path.join(HERE, "izhikevich_golden.npz"), **out)


def main():
	if os.environ.get("SNN_GOLDEN_ONLY") == "izhikevich":
		izhikevich_fixture()
		return
	encoder_fixture()
	dynamics_fixture()
	init_fixture()
	stacked_fixture()
	izhikevich_fixture()
	for f in sorted(os.listdir(HERE)):
		if f.endswith(".npz"):
			print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KB")


def stacked_fixture():
	"""Two stacked hidden layers (snn.py:116-128; train.py:72 shows the list form): state_dict, traces, loss, grads."""
	out = {}
	names = []
	for name, layer, widths, rec in (("lif_32_64", LayerType.LIF, [32, 64], True), ("alif_64_32", LayerType.ALIF, [64, 32], True),
			("alif_100_32_nonrec", LayerType.ALIF, [100, 32], False)):
		torch.manual_seed(11)
		kw = dict(learn_beta=False) if layer == LayerType.ALIF else {}
		net = SNN(48, 10, widths, use_recurrent_connection=rec, int_time_steps=14, spike_func=SpikeFuncType.FastSigmoid,
			hidden_layer_type=layer, device=torch.device("cpu"), **kw)
		g = torch.Generator().manual_seed(3)
		x = (torch.rand(4, 14, 48, generator=g) < (0.3 if layer == LayerType.ALIF else 0.12)).float()
		labels = torch.randint(0, 10, (4,), generator=g)
		net.train()
		logp, y, hs = net.get_prediction_log_proba(x, re_outputs_trace=True, re_hidden_states=True)
		loss = torch.nn.NLLLoss()(logp, labels)
		net.zero_grad()
		loss.backward()
		out[f"{name}/x"] = x.numpy().astype(np.uint8)
		out[f"{name}/labels"] = labels.numpy()
		out[f"{name}/y"] = y.detach().numpy()
		out[f"{name}/loss"] = np.float32(loss.item())
		out[f"{name}/widths"] = np.array(widths)
		out[f"{name}/flags"] = np.array([int(layer == LayerType.ALIF), int(rec)])
		for lname, tr in hs.items():
			out[f"{name}/Z/{lname}"] = tr[-1].detach().numpy().astype(np.float32)
			out[f"{name}/V/{lname}"] = tr[0].detach().numpy()
		keys = []
		for k, v in net.state_dict().items():
			out[f"{name}/sd/{k}"] = v.numpy()
			keys.append(k)
		out[f"{name}/keys"] = np.array(keys)
		for k, p in net.named_parameters():
			if p.grad is not None:
				out[f"{name}/grad/{k}"] = p.grad.numpy()
		names.append(name)
		print(name, "mean rates", {k: float(v[-1].mean()) for k, v in hs.items() if k != "readout"})
	out["names"] = np.array(names)
	np.savez_compressed(os.path.join(HERE, "stacked_golden.npz"), **out)


if __name__ == "__main__":
	if os.environ.get("SNN_GOLDEN_ONLY") == "stacked":
		stacked_fixture()
	else:
		main()
