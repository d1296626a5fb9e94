"""PyTorch-CPU restatement of the reference's hot path -- TEST INFRASTRUCTURE ONLY.

This is the algorithm exactly as the reference executes it: a Python loop over
time steps issuing small ``torch.matmul`` + elementwise ops, differentiated by
autograd with a hand-written surrogate backward.  It exists because the
reference itself (pure Python under /root/reference) cannot travel to the GPU
box; it is pinned against the reference by tests/test_oracle_golden.py using
fixtures the reference produced in the build container.

Used by: tests (second checker beside the C oracle) and bench.py's
``cpu_baseline`` / ``--impl reference`` legs (the timed CPU baseline, with all
host threads torch can use).  Never imported by the product package.

Citations are into the reference checkout.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch


class _FastSigmoidSpike(torch.autograd.Function):
	"""spike_funcs.py:12-29 forward, :46-62 backward."""

	@staticmethod
	def forward(ctx, v, thr, gamma):
		ctx.save_for_backward(v, thr, gamma)
		return (v >= thr).to(v.dtype)

	@staticmethod
	def backward(ctx, g):
		v, thr, gamma = ctx.saved_tensors
		return g / (gamma * (v - thr).abs() + 1.0) ** 2, None, None


class _PhiSpike(torch.autograd.Function):
	"""spike_funcs.py:12-29 forward, :65-79 backward (epsilon = 1e-5)."""

	@staticmethod
	def forward(ctx, v, thr, gamma):
		ctx.save_for_backward(v, thr, gamma)
		return (v >= thr).to(v.dtype)

	@staticmethod
	def backward(ctx, g):
		v, thr, gamma = ctx.saved_tensors
		te = thr + 1e-5
		return g * (gamma / te) * torch.clamp(1 - ((v - thr) / te).abs(), min=0.0), None, None


SURROGATES = {0: _FastSigmoidSpike, 1: _PhiSpike}


class TorchPortSNN:
	"""One hidden LIF (0) / ALIF (1) / Izhikevich (2) layer + leaky readout (snn.py:201-219)."""

	def __init__(
			self, N: int, H: int, O: int, T: int, layer_type: int = 1, surrogate: int = 0, recurrent: bool = True,
			dt: float = 1e-3, learn_beta: bool = False, seed: Optional[int] = None, **kw,
	):
		self.N, self.H, self.O, self.T = N, H, O, T
		self.layer_type, self.surrogate, self.recurrent = layer_type, surrogate, bool(recurrent)
		# spiking_layers.py:124-130 / :201-210 / :380-381 defaults
		self.dt = dt
		# IzhikevichLayer constants, spiking_layers.py:287-300 (gamma 1.0: the never-true isinstance test)
		self.iz = {k: torch.tensor(kw.get(k, v), dtype=torch.float32) for k, v in dict(
			C=100.0, v_rest=-60.0, v_th=-40.0, k=0.7, a=0.03, b=-2.0, c=-50.0, d=100.0, v_peak=35.0).items()}
		if layer_type == 2:
			tau_m, theta, gamma = 10.0 * dt, 1.0, kw.get("gamma", 1.0)      # weights ~ N(0, 1) (spiking_layers.py:302-307)
		elif layer_type == 0:
			tau_m, theta, gamma = kw.get("tau_m", 10.0 * dt), kw.get("threshold", 1.0), kw.get("gamma", 1.0)
		else:
			tau_m, theta, gamma = kw.get("tau_m", 20.0 * dt), kw.get("threshold", 0.03), kw.get("gamma", 0.3)
		tau_a, beta, tau_out = kw.get("tau_a", 200.0 * dt), kw.get("beta", 1.6), kw.get("tau_out", 10.0 * dt)
		f32 = torch.float32
		self.alpha = torch.tensor(math.exp(-dt / tau_m), dtype=f32)
		self.rho = torch.tensor(math.exp(-dt / tau_a), dtype=f32)
		self.kappa = torch.tensor(math.exp(-dt / tau_out), dtype=f32)
		self.theta = torch.tensor(theta, dtype=f32)
		self.gamma = torch.tensor(gamma, dtype=f32)
		self.beta = torch.tensor(beta, dtype=f32)
		g = torch.Generator().manual_seed(seed) if seed is not None else None
		self.W_in = (torch.randn(N, H, generator=g) * theta).requires_grad_()
		self.W_rec = (torch.randn(H, H, generator=g) * theta).requires_grad_() if recurrent else None
		self.rec_mask = 1.0 - torch.eye(H) if recurrent else None
		self.W_out = torch.randn(H, O, generator=g).requires_grad_()
		self.b_out = torch.zeros(O).requires_grad_()
		if learn_beta and layer_type == 1:
			# reference quirk: a learnable beta is re-drawn ~ N(0, theta^2) and never gets a gradient
			self.beta = (torch.randn((), generator=g) * theta).requires_grad_()

	def parameters(self):
		ps = [self.W_in]
		if self.W_rec is not None:
			ps.append(self.W_rec)
		if self.beta.requires_grad:
			ps.append(self.beta)
		return ps + [self.W_out, self.b_out]

	def load(self, W_in, W_rec, W_out, b_out, beta=None):
		with torch.no_grad():
			self.W_in.copy_(torch.as_tensor(W_in))
			if self.W_rec is not None:
				self.W_rec.copy_(torch.as_tensor(W_rec))
			self.W_out.copy_(torch.as_tensor(W_out))
			self.b_out.copy_(torch.as_tensor(b_out))
			if beta is not None:
				self.beta.copy_(torch.as_tensor(beta, dtype=torch.float32))

	def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, Tuple[torch.Tensor, ...]]]:
		B = x.shape[0]
		x = x.float()
		spike = SURROGATES[self.surrogate].apply
		V = torch.zeros(B, self.H, requires_grad=True)
		if self.layer_type == 2:
			V = (self.iz["v_rest"] * torch.ones(B, self.H)).requires_grad_()       # spiking_layers.py:309
		a = torch.zeros(B, self.H, requires_grad=True)
		Z = torch.zeros(B, self.H, requires_grad=True)
		y = torch.zeros(B, self.O, requires_grad=True)
		Vs, As, Zs, ys = [], [], [], []
		for t in range(self.T):
			cur = torch.matmul(x[:, t], self.W_in)                               # spiking_layers.py:163/233
			rec = torch.matmul(Z, self.W_rec * self.rec_mask) if self.recurrent else 0.0  # :165/235
			if self.layer_type == 2:                                                # spiking_layers.py:343-349
				z, r = self.iz, Z.detach()
				I = cur + rec
				dVdt = z["k"] * (V - z["v_rest"]) * (V - z["v_th"]) - a + I
				nV = (V + self.dt * dVdt / z["C"]) * (1.0 - r) + z["c"] * r
				a = (a + self.dt * (z["a"] * (z["b"] * (V - z["v_rest"]) - a))) + z["d"] * r
				V = nV
				Z = spike(V, z["v_peak"], self.gamma)
				y = self.kappa * y + torch.matmul(Z, self.W_out) + self.b_out
				Vs.append(V); As.append(a); Zs.append(Z); ys.append(y)
				continue
			V = (self.alpha * V + cur + rec) * (1.0 - Z.detach())                  # :169/239
			if self.layer_type == 1:
				a = self.rho * a + Z                                                # :240
				thr = self.theta + self.beta * a                                    # :241
			else:
				thr = self.theta
			Z = spike(V, thr, self.gamma)                                           # :170/242
			y = self.kappa * y + torch.matmul(Z, self.W_out) + self.b_out          # :407
			Vs.append(V); As.append(a); Zs.append(Z); ys.append(y)
		st = lambda l: torch.stack(l, dim=1)  # noqa: E731  (snn.py:195-199, :218)
		hidden = (st(Vs), st(As), st(Zs)) if self.layer_type != 0 else (st(Vs), st(Zs))
		out = st(ys)
		return out, {"input": hidden, "readout": (out,)}

	def log_proba(self, x):
		out, hs = self.forward(x)
		logits, _ = torch.max(out, dim=1)                                          # snn.py:228
		return torch.log_softmax(logits, dim=-1), out, hs                         # snn.py:258

	def exec_batch(self, x, labels, optimizer=None) -> float:
		"""snn.py:384-415 in train mode: forward, NLL loss, backward, optimizer step."""
		logp, _, _ = self.log_proba(x)
		loss = torch.nn.functional.nll_loss(logp, labels.long())                  # snn.py:297, :410
		if optimizer is not None:
			optimizer.zero_grad()
		else:
			for p in self.parameters():
				p.grad = None
		loss.backward()                                                            # snn.py:413
		if optimizer is not None:
			optimizer.step()
		return loss.item()
