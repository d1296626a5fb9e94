"""torchrun worker of test_gpu_dp.py (one process per GPU over NCCL): the fused exchange+Adam step
(snnk_adam_step_dp) against the plain path (NCCL mean all-reduce, then snnk_adam_step), eagerly and in a CUDA graph."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
	rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
	dev = torch.device("cuda", local)
	torch.cuda.set_device(dev)
	dist.init_process_group("nccl", device_id=dev)
	from snnimageclassification_b200 import FusedAdam, LayerType, SNN, SpikeFuncType
	from snnimageclassification_b200.distributed import allreduce_mean_
	res = {}

	# (1) raw optimizer: random per-rank gradients, 5 steps
	g = torch.Generator().manual_seed(0)
	shapes = [(784, 128), (128, 128), (128, 10), (10,), (3, 5, 7)]
	base = [torch.randn(s, generator=g) for s in shapes]
	pa = [b.clone().to(dev).requires_grad_() for b in base]
	pb = [b.clone().to(dev).requires_grad_() for b in base]
	plain, fused = FusedAdam(pa, lr=1e-2, weight_decay=1e-5), FusedAdam(pb, lr=1e-2, weight_decay=1e-5)
	assert fused.enable_data_parallel()
	gr = torch.Generator().manual_seed(100 + rank)
	worst, worst_g = 0.0, 0.0
	for it in range(5):
		for a, b in zip(pa, pb):
			gg = torch.randn(a.shape, generator=gr).to(dev)
			a.grad, b.grad = gg.clone(), gg.clone()
		allreduce_mean_(p.grad for p in pa)
		plain.step()
		fused.step()
		for a, b in zip(pa, pb):
			worst = max(worst, float((a - b).abs().max() / a.abs().max()))
			worst_g = max(worst_g, float((a.grad - b.grad).abs().max()))
	res["optimizer_rel"], res["grad_abs"] = worst, worst_g
	# replicas stay bit-identical across ranks
	flat = torch.cat([p.detach().reshape(-1) for p in pb])
	gathered = [torch.empty_like(flat) for _ in range(world)]
	dist.all_gather(gathered, flat)
	res["replicas_identical"] = all(torch.equal(gathered[0], t) for t in gathered)

	# (2) the captured training step of SNN with the fused exchange == eager step with the NCCL all-reduce
	def make():
		torch.manual_seed(0)
		return SNN(64, 10, 128, use_recurrent_connection=True, int_time_steps=20, spike_func=SpikeFuncType.FastSigmoid,
			hidden_layer_type=LayerType.ALIF, device=dev, learn_beta=True)
	gx = torch.Generator().manual_seed(7 + rank)
	x = (torch.rand(32, 20, 64, generator=gx) < 0.2).float().to(dev)
	y = torch.randint(0, 10, (32,), generator=gx).to(dev)
	crit = torch.nn.NLLLoss()
	n1, n2 = make(), make()
	o1 = FusedAdam(n1.parameters(), lr=1e-3, weight_decay=1e-5)
	o2 = FusedAdam(n2.parameters(), lr=1e-3, weight_decay=1e-5)
	assert o2.enable_data_parallel()
	n1.train(); n2.train()
	step = n2.graphed_train_step(x, y, crit, o2, static_inputs=True)
	for it in range(4):
		loss = n1.batch_loss(x, y, crit)
		o1.zero_grad()
		loss.backward()
		n1._allreduce_gradients(o1)
		o1.step()
		step()
	torch.cuda.synchronize()
	res["snn_rel"] = max(float((a - b).abs().max() / a.abs().max().clamp_min(1e-30))
		for a, b in zip(n1.parameters(), n2.parameters()))
	allres = [None] * world
	dist.all_gather_object(allres, res)
	if rank == 0:
		print("DPRESULT " + json.dumps(allres), flush=True)
	torch.cuda.synchronize()
	os._exit(0)


if __name__ == "__main__":
	main()
