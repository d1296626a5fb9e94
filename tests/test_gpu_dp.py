"""Data-parallel optimizer step on the GPU: snnk_adam_step_dp (gradient exchange over NVLink peer memory + mean +
Adam in one kernel) against snnk_adam_step, and -- on a box with at least two GPUs -- against the NCCL all-reduce."""
import ctypes
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _arr(ts):
	return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def test_adam_dp_single_rank_is_plain_adam():
	"""world = 1: the exchange degenerates to a copy through the local slot; the update must be bit-identical to
	snnk_adam_step, over several launches (epoch parity alternates) and a tensor count that leaves ragged tails."""
	from snnimageclassification_b200 import _cabi
	lib = _cabi.lib()
	g = torch.Generator().manual_seed(3)
	shapes = [(784, 128), (128, 128), (128, 10), (10,), (1,), (37, 3)]
	P = [torch.randn(s, generator=g).to(DEV) for s in shapes]
	Q = [p.clone() for p in P]
	mk = lambda: ([torch.zeros_like(p) for p in P], [torch.zeros_like(p) for p in P],
		[torch.zeros((), device=DEV) for _ in P])
	(m1, v1, s1), (m2, v2, s2) = mk(), mk()
	numel = (ctypes.c_int64 * len(P))(*[p.numel() for p in P])
	total = sum(p.numel() for p in P)
	nbytes = ctypes.c_size_t(0)
	_cabi.check(lib.snnk_adam_dp_buffer_bytes(1, total, ctypes.byref(nbytes)), "bytes")
	assert nbytes.value == 2 * total * 8
	xbuf = torch.zeros(nbytes.value // 4, dtype=torch.float32, device=DEV)
	state = torch.zeros(16, dtype=torch.int32, device=DEV)
	peers = (ctypes.c_void_p * 1)(xbuf.data_ptr())
	for it in range(5):
		G = [torch.randn(s, generator=g).to(DEV) for s in shapes]
		G2 = [x.clone() for x in G]
		_cabi.check(lib.snnk_adam_step(len(P), _arr(P), _arr(G), _arr(m1), _arr(v1), _arr(s1), numel, 1e-2, 0.9, 0.999,
			1e-8, 1e-5, _cabi.stream_ptr()), "adam")
		_cabi.check(lib.snnk_adam_step_dp(len(Q), _arr(Q), _arr(G2), _arr(m2), _arr(v2), _arr(s2), numel, 1e-2, 0.9, 0.999,
			1e-8, 1e-5, 0, 1, peers, state.data_ptr(), _cabi.stream_ptr()), "adam_dp")
		torch.cuda.synchronize()
		for a, b in zip(P, Q):
			assert torch.equal(a, b)
		for a, b in zip(G, G2):
			assert torch.equal(a, b)
	assert state[:4].tolist() == [5, 0, 0, 0]
	stamps = state[4:12].cpu().view(torch.int64).tolist()
	assert stamps[0] > 0 and stamps == sorted(stamps)
	assert all(float(s) == 5.0 for s in s2)
	# argument checks
	assert lib.snnk_adam_step_dp(len(Q), _arr(Q), _arr(G2), _arr(m2), _arr(v2), _arr(s2), numel, 1e-2, 0.9, 0.999, 1e-8,
		1e-5, 1, 1, peers, state.data_ptr(), _cabi.stream_ptr()) != 0
	assert lib.snnk_adam_step_dp(len(Q), _arr(Q), _arr(G2), _arr(m2), _arr(v2), _arr(s2), numel, 1e-2, 0.9, 0.999, 1e-8,
		1e-5, 0, 17, peers, state.data_ptr(), _cabi.stream_ptr()) != 0


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_exchange_matches_nccl_allreduce():
	n = min(torch.cuda.device_count(), 8)
	worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dp_nccl_worker.py")
	cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
		"--master-port", str(_free_port()), worker]
	out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
	assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
	line = [l for l in out.stdout.splitlines() if l.startswith("DPRESULT ")][-1]
	for res in json.loads(line[len("DPRESULT "):]):
		assert res["replicas_identical"]
		assert res["optimizer_rel"] <= 1e-6 and res["grad_abs"] <= 1e-6
		assert res["snn_rel"] <= 1e-5
