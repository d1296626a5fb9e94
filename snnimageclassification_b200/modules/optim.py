"""``FusedAdam`` -- the reference's default optimizer (``torch.optim.Adam(lr, weight_decay=1e-5)``, snn.py:299) with
its ``step()`` running as ONE launch of ``snnk_adam_step`` over all parameter tensors.

It IS a ``torch.optim.Adam`` (same constructor, ``state_dict`` layout, per-parameter ``step``/``exp_avg``/
``exp_avg_sq`` state), so the reference's checkpoints load into it and its own load into ``torch.optim.Adam``;
only the arithmetic of ``step`` moves to libsnnk.  The step counters live on the device, which lets the whole
training step sit inside one captured CUDA graph (modules/graphed.py).
"""
from __future__ import annotations

import ctypes

import torch

from .. import _cabi


class FusedAdam(torch.optim.Adam):
	def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
		super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True, foreach=False)

	@torch.no_grad()
	def step(self, closure=None):
		loss = None
		if closure is not None:
			with torch.enable_grad():
				loss = closure()
		lib = _cabi.lib()
		for group in self.param_groups:
			ps, gs, ms, vs, ss = [], [], [], [], []
			for p in group["params"]:
				if p.grad is None:
					continue
				if not p.is_cuda:
					raise RuntimeError("FusedAdam runs on CUDA sm_100 parameters only; there is no CPU fallback")
				st = self.state[p]
				if len(st) == 0:   # same lazy initialisation as torch.optim.Adam with capturable=True
					st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
					st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
					st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
				g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
				ps.append(p); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"]); ss.append(st["step"])
			lr = float(group["lr"]) if not torch.is_tensor(group["lr"]) else float(group["lr"].item())
			b1, b2 = group["betas"]
			for k0 in range(0, len(ps), 16):
				chunk = slice(k0, k0 + 16)
				n = len(ps[chunk])
				arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
				numel = (ctypes.c_int64 * n)(*[p.numel() for p in ps[chunk]])
				with torch.cuda.device(ps[k0].device):
					rc = lib.snnk_adam_step(
						n, arr(ps[chunk]), arr(gs[chunk]), arr(ms[chunk]), arr(vs[chunk]), arr(ss[chunk]), numel, lr, b1, b2,
						float(group["eps"]), float(group["weight_decay"]), _cabi.stream_ptr())
				_cabi.check(rc, "snnk_adam_step")
		return loss
