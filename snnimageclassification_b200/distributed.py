"""Data-parallel plumbing: one process per GPU, batch rows sharded, weights replicated.

The spiking path has exactly one exchange step per training iteration: the mean of the (small) weight gradients
over the ranks.  With ``FusedAdam.enable_data_parallel()`` it happens inside the optimizer kernel over NVLink peer
memory (``snnk_adam_step_dp``; ``PeerExchangeBuffer`` below supplies the peer-mapped buffers); for any other optimizer it
is ONE coalesced all-reduce (NCCL on GPUs, gloo in the CPU tests).  With equal shards and a mean-reduced local loss this
reproduces the single-process full-batch gradient up to summation order.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


def world_size() -> int:
	return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class PeerExchangeBuffer:
	"""``nbytes`` of zero-filled device memory on every rank of ``group``, each rank's block mapped into every other
	rank's address space (NVLink peer access through torch's symmetric-memory allocator -- plumbing only: the
	kernels that read and write it are libsnnk's).  ``ptrs[r]`` is rank r's block as seen from this process."""

	def __init__(self, nbytes: int, device: torch.device, group=None):
		import torch.distributed._symmetric_memory as symm
		group = group if group is not None else dist.group.WORLD
		self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
		self.buf = symm.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
		self.buf.zero_()
		torch.cuda.synchronize(device)
		self.handle = symm.rendezvous(self.buf, group.group_name)
		self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
		if len(self.ptrs) != self.world or self.ptrs[self.rank] != self.buf.data_ptr():
			raise RuntimeError("symmetric-memory rendezvous returned an unexpected peer table")
		dist.barrier(group)     # nobody pushes before every rank has zeroed its flags
		torch.cuda.synchronize(device)


def allreduce_mean_(tensors: Iterable[torch.Tensor]) -> None:
	"""In-place mean over the ranks of every tensor in ``tensors`` with a single flat all-reduce."""
	ws = world_size()
	if ws == 1:
		return
	ts: List[torch.Tensor] = [t for t in tensors if t is not None]
	if not ts:
		return
	if ts[0].is_cuda and dist.get_backend() == "nccl" and hasattr(dist, "_coalescing_manager"):
		# NCCL: the per-tensor all-reduces are coalesced into ONE group launch that averages in place -- no flatten /
		# copy-back kernels around the collective (it all sits inside the captured training graph)
		try:
			with dist._coalescing_manager(device=ts[0].device):
				for t in ts:
					dist.all_reduce(t, op=dist.ReduceOp.AVG)
			return
		except Exception:   # pragma: no cover  (older torch: fall through to the flat buffer)
			pass
	flat = torch.cat([t.reshape(-1) for t in ts])
	dist.all_reduce(flat, op=dist.ReduceOp.SUM)
	flat.div_(ws)
	off = 0
	for t in ts:
		n = t.numel()
		t.copy_(flat[off:off + n].view_as(t))
		off += n


def shard_batch(x: torch.Tensor, rank: int, ws: int) -> torch.Tensor:
	"""Equal contiguous shard of the batch dimension (equal shards are required for exact equivalence)."""
	assert x.shape[0] % ws == 0, "global batch must be divisible by the number of ranks"
	n = x.shape[0] // ws
	return x[rank * n:(rank + 1) * n]
