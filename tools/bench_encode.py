"""Times ToSpikes.encode_batch variants as CUDA-graph replays (device time per call, us)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snnimageclassification_b200 import ToSpikes

dev = torch.device("cuda:0")
B, N, T = 256, 784, 100
g = torch.Generator().manual_seed(0)
k = torch.randint(1, 256, (B, N), generator=g).float() / 255.0
img = torch.where(torch.rand(B, N, generator=g) < 0.19, k, torch.zeros(())).to(dev)
for tau in (0.02, 20.0):
	enc = ToSpikes(T, tau=tau, use_periods=True)
	for name, kw in (("dense, no runs", dict(frame_runs=False)), ("runs", dict()), ("runs, lazy", dict(lazy=True))):
		for _ in range(3):
			enc.encode_batch(img, **kw)
		torch.cuda.synchronize()
		gr = torch.cuda.CUDAGraph()
		with torch.cuda.graph(gr):
			out = enc.encode_batch(img, **kw)
		for _ in range(5):
			gr.replay()
		torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		for _ in range(200):
			gr.replay()
		e1.record()
		torch.cuda.synchronize()
		print(f"tau={tau:5}  {name:16s} {e0.elapsed_time(e1) / 200 * 1e3:8.2f} us")
