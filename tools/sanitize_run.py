#!/usr/bin/env python
"""Small-shape tour of every kernel family of libsnnk.so for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

One tool per gpurun call (profiling guide).  Shapes are tiny (racecheck slows kernels ~100x) but reach every code
path with barriers / mbarrier rings / aliased staging / peer-protocol words: K5 encoder (+ run table, lazy rows, bit
raster), K1/K4 tcgen05 GEMMs (dense, compact, Z-only), SIMT GEMMs, K2/K3 SIMT recurrences (H = 32/64/128, R = 1),
the tensor-core recurrences (recur_tc.cuh; SNNK_MMA_RECUR=0 for the SIMT ones in that mode), the wide path, K6 head, Adam and the data-parallel Adam with
world = 1.  Prints one line per stage; exits non-zero if a result is not finite.
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from snnimageclassification_b200 import FusedAdam, LayerType, SNN, SpikeFuncType, ToSpikes, _cabi  # noqa: E402
from snnimageclassification_b200.modules import functional as F_  # noqa: E402

DEV = torch.device("cuda:0")


def train_step(net, x, y, tag):
	net.train()
	net.zero_grad()
	loss = net.batch_loss(x, y)
	loss.backward()
	torch.cuda.synchronize()
	ok = bool(torch.isfinite(loss)) and all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)
	print(f"{tag}: loss {loss.item():.5f} finite={ok}", flush=True)
	if not ok:
		sys.exit(3)


def main():
	_cabi.require_b200(DEV)
	g = torch.Generator().manual_seed(0)
	T, N, O = 10, 64, 10
	img = (torch.randint(1, 256, (12, N), generator=g).float() / 255.0) * (torch.rand(12, N, generator=g) < 0.3)
	lab = torch.randint(0, O, (12,), generator=g).to(DEV)
	# encoder: dense raster, run table (eager + lazy), bit raster, tau = 20 regime
	for periodic in (False, True):
		enc = ToSpikes(T, use_periods=periodic)
		x = enc.encode_batch(img.to(DEV))
		xl = enc.encode_batch(img.to(DEV), lazy=True)
		bits = enc.encode_batch_bits(img.to(DEV))
		torch.cuda.synchronize()
		print(f"encode periodic={periodic}: {float(x.mean()):.4f} runs={int(F_.get_runs(x)[0])} bits={tuple(bits.shape)}", flush=True)
	x20 = ToSpikes(T, use_periods=False, tau=20.0).encode_batch(img.to(DEV))
	xp = ToSpikes(T, use_periods=True).encode_batch(img.to(DEV))

	for H in (32, 64, 128):
		for tc in (False, True):
			for layer, rec in ((LayerType.ALIF, True), (LayerType.LIF, False)):
				torch.manual_seed(1)
				net = SNN(N, O, H, use_recurrent_connection=rec, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
					hidden_layer_type=layer, device=DEV, tensor_core=tc, **({"learn_beta": True} if layer == LayerType.ALIF else {}))
				train_step(net, xp, lab, f"H={H} tc={tc} {layer.name} rec={rec} dedup-input")
				train_step(net, x20, lab, f"H={H} tc={tc} {layer.name} rec={rec} dense-input")
	# Phi surrogate, Izhikevich, stacked layers
	torch.manual_seed(2)
	train_step(SNN(N, O, 64, int_time_steps=T, spike_func=SpikeFuncType.Phi, hidden_layer_type=LayerType.ALIF, device=DEV,
		learn_beta=True), x20, lab, "Phi ALIF H=64")
	train_step(SNN(N, O, 32, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid, hidden_layer_type=LayerType.Izhikevich,
		device=DEV), x20, lab, "Izhikevich H=32")
	train_step(SNN(N, O, [32, 64], int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid, hidden_layer_type=LayerType.LIF,
		device=DEV), x20, lab, "stacked 32-64")
	# wide path
	torch.manual_seed(3)
	for H in (256, 1024):
		net = SNN(N, O, H, use_recurrent_connection=True, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
			hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True)
		train_step(net, x20, lab, f"wide H={H}")
		net.eval()
		with torch.no_grad():
			lg = net.get_prediction_logits(x20, re_outputs_trace=False, re_hidden_states=False)
		print(f"wide H={H} inference finite={bool(torch.isfinite(lg).all())}", flush=True)
	# the tensor-core recurrences (recur_tc.cuh) forced at a small, ragged batch (21 rows: a partial 8-row tile)
	img2 = (torch.randint(1, 256, (21, N), generator=g).float() / 255.0) * (torch.rand(21, N, generator=g) < 0.3)
	lab2 = torch.randint(0, O, (21,), generator=g).to(DEV)
	torch.manual_seed(4)
	net = SNN(N, O, 128, use_recurrent_connection=True, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True)
	os.environ["SNNK_MMA_RECUR"] = "1"      # the library picks them from B >= 1024 on its own
	train_step(net, ToSpikes(T, use_periods=True).encode_batch(img2.to(DEV)), lab2, "tensor-core recurrence B=21, dedup input")
	train_step(net, ToSpikes(T, use_periods=False, tau=20.0).encode_batch(img2.to(DEV)), lab2, "tensor-core recurrence B=21, dense input")
	os.environ.pop("SNNK_MMA_RECUR")

	# optimizer: plain and data-parallel form with world = 1 (its own buffer is the only peer)
	opt = FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
	opt.step()
	ps = [p for p in net.parameters() if p.grad is not None]
	total = sum(p.numel() for p in ps)
	nbytes = ctypes.c_size_t(0)
	lib = _cabi.lib()
	_cabi.check(lib.snnk_adam_dp_buffer_bytes(1, total, ctypes.byref(nbytes)), "buffer_bytes")
	buf = torch.zeros(nbytes.value, dtype=torch.uint8, device=DEV)
	state = torch.zeros(16, dtype=torch.int32, device=DEV)
	n = len(ps)
	arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
	sts = [opt.state[p] for p in ps]
	for _ in range(2):
		rc = lib.snnk_adam_step_dp(n, arr(ps), arr([p.grad for p in ps]), arr([s["exp_avg"] for s in sts]),
			arr([s["exp_avg_sq"] for s in sts]), arr([s["step"] for s in sts]), (ctypes.c_int64 * n)(*[p.numel() for p in ps]),
			1e-3, 0.9, 0.999, 1e-8, 1e-5, 0, 1, (ctypes.c_void_p * 1)(buf.data_ptr()), state.data_ptr(), _cabi.stream_ptr())
		_cabi.check(rc, "snnk_adam_step_dp")
	torch.cuda.synchronize()
	print(f"adam + adam_dp(world=1): finite={all(bool(torch.isfinite(p).all()) for p in ps)}", flush=True)
	print("SANITIZE_TOUR_DONE", flush=True)


if __name__ == "__main__":
	main()
