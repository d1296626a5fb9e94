"""Data-parallel path on CPU: world_size 2 over gloo (the N > 1 path of bench.py uses the same code over NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import rel_err


def _free_port():
	with socket.socket() as s:
		s.bind(("127.0.0.1", 0))
		return s.getsockname()[1]


def _worker(rank, ws, port, out):
	os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
	dist.init_process_group("gloo", rank=rank, world_size=ws)
	torch.set_num_threads(1)
	from snnimageclassification_b200 import LayerType, SNN, SpikeFuncType
	from snnimageclassification_b200.distributed import shard_batch
	from oracle.torch_port import TorchPortSNN
	# (1) SNN._allreduce_gradients: one flat all-reduce, mean over ranks, None grads (beta) skipped
	torch.manual_seed(0)
	net = SNN(16, 10, 32, hidden_layer_type=LayerType.ALIF, spike_func=SpikeFuncType.FastSigmoid,
		device=torch.device("cpu"), learn_beta=True, int_time_steps=4)
	for p in net.parameters():
		p.grad = torch.full_like(p, float(rank + 1))
	net.layers["input"].beta.grad = None
	net._allreduce_gradients()
	ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in net.parameters() if p.grad is not None)
	ok = ok and net.layers["input"].beta.grad is None
	# (2) sharded batch + gradient mean == full-batch gradient (local losses are means over equal shards)
	g = torch.Generator().manual_seed(5)
	x = (torch.rand(8, 6, 16, generator=g) < 0.3).float()
	y = torch.randint(0, 10, (8,), generator=g)
	full = TorchPortSNN(16, 32, 10, 6, layer_type=1, recurrent=True, seed=1)
	full.exec_batch(x, y)
	local = TorchPortSNN(16, 32, 10, 6, layer_type=1, recurrent=True, seed=1)
	local.exec_batch(shard_batch(x, rank, ws), shard_batch(y, rank, ws))
	from snnimageclassification_b200.distributed import allreduce_mean_
	allreduce_mean_(p.grad for p in local.parameters())
	err = max(rel_err(a.grad.numpy(), b.grad.numpy()) for a, b in zip(local.parameters(), full.parameters()))
	out[rank] = (ok, err)
	dist.destroy_process_group()


def test_dp_gradient_mean_world_size_2():
	ctx = mp.get_context("spawn")
	with ctx.Manager() as m:
		out = m.dict()
		port = _free_port()
		procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
		for p in procs:
			p.start()
		for p in procs:
			p.join(120)
			assert p.exitcode == 0
		for r in range(2):
			ok, err = out[r]
			assert ok, f"rank {r}: flat all-reduce mean wrong"
			assert err <= 1e-5, f"rank {r}: sharded gradient differs from the full-batch gradient by {err:.2e}"
