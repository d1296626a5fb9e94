#!/usr/bin/env python
"""A few EAGER training steps of one BASELINE configuration, for ncu (kernels appear as plain launches):

    ncu --set full --clock-control none --import-source on -k regex:<pattern> -s <skip> -c <n> -o gpurun_out/prof \
        python tools/profile_step.py [--config c2|c1|c3|c4] [--batch B] [--steps K] [--dense] [--infer]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snnimageclassification_b200 import FusedAdam, LayerType, SNN, SpikeFuncType, ToSpikes  # noqa: E402

CFG = {"c1": (128, LayerType.LIF, False, 0.19, 256), "c2": (128, LayerType.ALIF, True, 0.19, 256),
	"c3": (64, LayerType.ALIF, False, 0.5, 256), "c4": (1024, LayerType.ALIF, True, 0.19, 512)}


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--config", default="c2")
	ap.add_argument("--batch", type=int, default=0)
	ap.add_argument("--steps", type=int, default=3)
	ap.add_argument("--T", type=int, default=100)
	ap.add_argument("--dense", action="store_true", help="drop the run table: dense kernels")
	ap.add_argument("--infer", action="store_true", help="no-trace inference instead of training steps")
	ap.add_argument("--bits", action="store_true", help="feed the bit-packed raster (SNNK_F_INPUT_BITS kernels)")
	ap.add_argument("--time", action="store_true", help="print the library's per-kernel CUDA-event times (eager launches)")
	a = ap.parse_args()
	H, layer, rec, ink, B = CFG[a.config]
	B = a.batch or B
	dev = torch.device("cuda:0")
	torch.manual_seed(0)
	net = SNN(784, 10, H, use_recurrent_connection=rec, int_time_steps=a.T, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=layer, device=dev, **({"learn_beta": True} if layer == LayerType.ALIF else {}))
	opt = FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
	g = torch.Generator().manual_seed(1)
	img = (torch.randint(1, 256, (B, 784), generator=g).float() / 255.0) * (torch.rand(B, 784, generator=g) < ink)
	lab = torch.randint(0, 10, (B,), generator=g).to(dev)
	enc = ToSpikes(a.T, use_periods=True)
	x = enc.encode_batch_bits(img.to(dev)) if a.bits else enc.encode_batch(img.to(dev), frame_runs=not a.dense)
	if a.infer:
		net.eval()
		with torch.no_grad():
			for _ in range(a.steps):
				out = net.get_prediction_logits(x, re_outputs_trace=False, re_hidden_states=False)
		torch.cuda.synchronize()
		print("logits", float(out.abs().mean()))
		if a.time:
			from snnimageclassification_b200 import _cabi
			with torch.no_grad(), _cabi.kernel_profile() as prof:
				for _ in range(10):
					net.get_prediction_logits(x, re_outputs_trace=False, re_hidden_states=False)
				torch.cuda.synchronize()
			for name, (ms, n) in prof.result.items():
				print(f"  {name:55s} {1e3 * ms / 10:9.1f} us/call  ({n // 10} launches)")
		return
	net.train()

	def steps(n):
		for _ in range(n):
			loss = net.batch_loss(x, lab)
			opt.zero_grad()
			loss.backward()
			opt.step()
		return loss
	loss = steps(a.steps)
	torch.cuda.synchronize()
	print("loss", loss.item())
	if a.time:
		from snnimageclassification_b200 import _cabi
		with _cabi.kernel_profile() as prof:
			steps(20)
			torch.cuda.synchronize()
		for name, (ms, n) in prof.result.items():
			print(f"  {name:55s} {1e3 * ms / 20:9.1f} us/step  ({n // 20} launches)")


if __name__ == "__main__":
	main()
