"""``FusedAdam`` -- the reference's default optimizer (``torch.optim.Adam(lr, weight_decay=1e-5)``, snn.py:299) with
its ``step()`` running as ONE launch of ``snnk_adam_step`` over all parameter tensors.

It IS a ``torch.optim.Adam`` (same constructor, ``state_dict`` layout, per-parameter ``step``/``exp_avg``/
``exp_avg_sq`` state), so the reference's checkpoints load into it and its own load into ``torch.optim.Adam``;
only the arithmetic of ``step`` moves to libsnnk.  The step counters live on the device, which lets the whole
training step sit inside one captured CUDA graph (modules/graphed.py).

Data parallel (one process per GPU): after ``enable_data_parallel()`` the same launch also performs
the path's only exchange step -- the mean of the weight gradients over the ranks -- by pushing the gradients through
NVLink peer memory (``snnk_adam_step_dp``); ``reduces_gradients`` then tells ``SNN._exec_batch`` to skip its
all-reduce.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _cabi


class FusedAdam(torch.optim.Adam):
	def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
		super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True, foreach=False)
		self.reduces_gradients = False
		self._dp_group = None
		self._dp_ctx = {}
		self._state_version = 0      # bumped whenever the state tensors are replaced (captured graphs hold their pointers)
		self._lr_cache = {}          # python value of a tensor lr, read outside graph capture by hyper_signature()

	def load_state_dict(self, state_dict):
		"""Accepts the state of a plain ``torch.optim.Adam`` / a reference checkpoint (snn.py:443-448) as well as its
		own.  torch's loader copies the SAVED group options over this optimizer's (``capturable=False``,
		``foreach=None``, ...) and leaves ``state['step']`` wherever ``torch.load`` put it (a CPU tensor under
		``map_location='cpu'``, an int in very old files); ``step()`` hands the counters' device pointers to the
		kernel, so they are normalised here: capturable groups, float32 0-d counters and moments on the parameter's
		device."""
		super().load_state_dict(state_dict)
		self._state_version += 1
		for group in self.param_groups:
			group["capturable"], group["foreach"], group["fused"] = True, False, None
			group.setdefault("maximize", False)
			group.setdefault("amsgrad", False)
			if group.get("maximize") or group.get("amsgrad"):
				raise RuntimeError("FusedAdam implements torch.optim.Adam without amsgrad / maximize only")
			for p in group["params"]:
				st = self.state.get(p)
				if not st:
					continue
				step = st.get("step", 0.0)
				step = step.detach().to(device=p.device, dtype=torch.float32) if torch.is_tensor(step) else torch.tensor(
					float(step), dtype=torch.float32, device=p.device)
				st["step"] = step.reshape(()).clone()
				for k in ("exp_avg", "exp_avg_sq"):
					if k in st:
						st[k] = st[k].detach().to(device=p.device, dtype=torch.float32).contiguous()

	def hyper_signature(self):
		"""The scalar hyper-parameters ``step()`` passes to the kernel BY VALUE.  A captured CUDA graph freezes them, so
		``GraphedTrainStep`` compares this signature before every replay and re-captures when a scheduler (or the
		user) has changed ``param_groups`` in between."""
		sig = [self._state_version]
		for gi, g in enumerate(self.param_groups):
			lr = g["lr"]
			if torch.is_tensor(lr):      # a host read: legal here (never called while capturing), not inside step()
				self._lr_cache[gi] = float(lr.item())
				lr = self._lr_cache[gi]
			sig.append((float(lr), tuple(float(b) for b in g["betas"]), float(g["eps"]), float(g["weight_decay"])))
		return tuple(sig)

	def enable_data_parallel(self, group=None) -> bool:
		"""Fuse the cross-rank gradient mean into ``step()``.  Returns False (and leaves the optimizer as it was:
		the caller keeps using the NCCL all-reduce) when there is a single rank, the backend is not NCCL or peer
		memory cannot be mapped.  Every rank must call it, and ``step()``, the same number of times."""
		import torch.distributed as dist
		if not (dist.is_available() and dist.is_initialized()):
			return False
		group = group if group is not None else dist.group.WORLD
		if dist.get_world_size(group) == 1 or dist.get_backend(group) != "nccl":
			return False
		if dist.get_world_size(group) > 16:
			return False
		# Probe the peer mapping once, and agree on the outcome: either every rank exchanges through peer memory or every
		# rank keeps the all-reduce (a mixed group would deadlock).
		ok = 1
		try:
			from ..distributed import PeerExchangeBuffer
			dev = next(p.device for g in self.param_groups for p in g["params"])
			if dev.type != "cuda":
				raise RuntimeError("parameters are not on a CUDA device")
			PeerExchangeBuffer(256, dev, group)
		except Exception as e:   # no peer access, no symmetric-memory allocator, ...
			import warnings
			warnings.warn(f"FusedAdam: peer-memory gradient exchange unavailable ({e!r}); keeping the NCCL all-reduce")
			ok = 0
		try:
			flag = torch.tensor([ok], dtype=torch.int32, device=dev if ok else torch.device("cuda", torch.cuda.current_device()))
			dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
			ok = int(flag.item())
		except Exception:
			ok = 0
		if not ok:
			return False
		self._dp_group = group
		self.reduces_gradients = True
		return True

	def exchange_timeline(self):
		"""Per chunk, microseconds of the LAST data-parallel launch on this rank's device clock, relative to kernel
		start, as seen by the first thread: (stores issued, all peers' values seen, kernel done)."""
		out = []
		for xbuf, state, _ in self._dp_ctx.values():
			t = state[4:12].cpu().view(torch.int64).tolist()
			out.append(tuple((t[i] - t[0]) / 1e3 for i in (1, 2, 3)))
		return out

	def _dp_context(self, key, ps):
		ctx = self._dp_ctx.get(key)
		if ctx is None:
			from ..distributed import PeerExchangeBuffer
			if torch.cuda.is_current_stream_capturing():
				raise RuntimeError("FusedAdam: the first data-parallel step() must run outside CUDA graph capture")
			total = sum(p.numel() for p in ps)
			nbytes = ctypes.c_size_t(0)
			world = torch.distributed.get_world_size(self._dp_group)
			_cabi.check(_cabi.lib().snnk_adam_dp_buffer_bytes(world, total, ctypes.byref(nbytes)), "snnk_adam_dp_buffer_bytes")
			xbuf = PeerExchangeBuffer(nbytes.value, ps[0].device, self._dp_group)
			state = torch.zeros(16, dtype=torch.int32, device=ps[0].device)
			ctx = self._dp_ctx[key] = (xbuf, state, total)
		if ctx[2] != sum(p.numel() for p in ps):
			raise RuntimeError("FusedAdam: the set of parameters with gradients changed between data-parallel steps")
		return ctx

	@torch.no_grad()
	def step(self, closure=None):
		loss = None
		if closure is not None:
			with torch.enable_grad():
				loss = closure()
		lib = _cabi.lib()
		for gi, group in enumerate(self.param_groups):
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
				if self.reduces_gradients and not p.grad.is_contiguous():
					p.grad = p.grad.contiguous()
				g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
				ps.append(p); gs.append(g); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"]); ss.append(st["step"])
			lr = group["lr"]
			if torch.is_tensor(lr):
				if torch.cuda.is_current_stream_capturing():
					if gi not in self._lr_cache:
						raise RuntimeError("FusedAdam: a tensor lr must be read before graph capture (call hyper_signature())")
					lr = self._lr_cache[gi]
				else:
					lr = self._lr_cache[gi] = float(lr.item())
			lr = float(lr)
			b1, b2 = group["betas"]
			for k0 in range(0, len(ps), 16):
				chunk = slice(k0, k0 + 16)
				n = len(ps[chunk])
				arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
				numel = (ctypes.c_int64 * n)(*[p.numel() for p in ps[chunk]])
				with torch.cuda.device(ps[k0].device):
					if self.reduces_gradients:
						xbuf, state, _ = self._dp_context((gi, k0), ps[chunk])
						peers = (ctypes.c_void_p * xbuf.world)(*xbuf.ptrs)
						rc = lib.snnk_adam_step_dp(
							n, arr(ps[chunk]), arr(gs[chunk]), arr(ms[chunk]), arr(vs[chunk]), arr(ss[chunk]), numel, lr, b1,
							b2, float(group["eps"]), float(group["weight_decay"]), xbuf.rank, xbuf.world, peers,
							state.data_ptr(), _cabi.stream_ptr())
					else:
						rc = lib.snnk_adam_step(
							n, arr(ps[chunk]), arr(gs[chunk]), arr(ms[chunk]), arr(vs[chunk]), arr(ss[chunk]), numel, lr, b1, b2,
							float(group["eps"]), float(group["weight_decay"]), _cabi.stream_ptr())
				_cabi.check(rc, "snnk_adam_step")
		return loss
