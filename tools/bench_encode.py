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

# ---- pieces of the lazy pipeline, each as its own graph -------------------------------------------------------------------
import ctypes
from snnimageclassification_b200 import _cabi
lib = _cabi.lib()
enc = ToSpikes(T, tau=0.02, use_periods=True)
out = torch.empty((B, T, N), dtype=torch.float32, device=dev)
changed = torch.zeros((B, T), dtype=torch.uint8, device=dev)
table = torch.empty((lib.snnk_run_table_bytes(B, T) // 4,), dtype=torch.int32, device=dev)
x = enc.encode_batch(img)        # fills a valid table / flags for the stand-alone pieces
from snnimageclassification_b200.modules.functional import get_runs
table.copy_(get_runs(x))
flags = ((x[:, 1:] != x[:, :-1]).any(dim=2)).to(torch.uint8)
changed[:, 1:] = flags


def timed(name, fn):
	for _ in range(3):
		fn()
	torch.cuda.synchronize()
	gr = torch.cuda.CUDAGraph()
	with torch.cuda.graph(gr):
		fn()
	for _ in range(5):
		gr.replay()
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(200):
		gr.replay()
	e1.record()
	torch.cuda.synchronize()
	print(f"{name:28s} {e0.elapsed_time(e1) / 200 * 1e3:8.2f} us")


st = lambda: _cabi.stream_ptr()
timed("snnk_frame_runs", lambda: lib.snnk_frame_runs(B, T, _cabi.ptr(changed), _cabi.ptr(table), st()))
timed("memset flags", lambda: changed.zero_())
timed("snnk_encode (dense)", lambda: lib.snnk_encode(_cabi.ptr(img), 0, B, N, T, float(T), 0.02, 0.2, 1e-7, 1, _cabi.ptr(out), 0, None, st()))
timed("snnk_encode_runs lazy", lambda: lib.snnk_encode_runs(_cabi.ptr(img), 0, B, N, T, float(T), 0.02, 0.2, 1e-7, 1, _cabi.ptr(out), 0, None, _cabi.ptr(changed), _cabi.ptr(table), 1, st()))
timed("snnk_encode_runs eager", lambda: lib.snnk_encode_runs(_cabi.ptr(img), 0, B, N, T, float(T), 0.02, 0.2, 1e-7, 1, _cabi.ptr(out), 0, None, _cabi.ptr(changed), _cabi.ptr(table), 0, st()))
