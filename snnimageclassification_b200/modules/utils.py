"""Small host-side helpers used by ``SNN.fit`` -- counterpart of the reference's src/modules/utils.py.

Only what the training loop needs is here (the loss-history container and the nested-dict merge used for the
checkpoint index); plotting is optional and skipped when matplotlib is not installed.
"""
from __future__ import annotations

import collections.abc
from collections import defaultdict
from typing import Dict, List

import numpy as np
import torch


def batchwise_temporal_filter(x: torch.Tensor, decay: float = 0.9) -> torch.Tensor:
	"""sum_t decay^(T-1-t) x[:, t] for x (B, T, ...) (reference utils.py:11-25; not on the live path)."""
	T = x.shape[1]
	assert T >= 1
	w = torch.pow(decay, torch.arange(T - 1, -1, -1, dtype=torch.float32, device=x.device))
	return (x * w.view(1, T, *([1] * (x.ndim - 2)))).sum(dim=1)


def mapping_update_recursively(d, u):
	"""Nested dict merge (reference utils.py:28-40)."""
	for k, v in u.items():
		if isinstance(v, collections.abc.Mapping):
			d[k] = mapping_update_recursively(d.get(k, {}), v)
		else:
			d[k] = v
	return d


class LossHistory:
	"""{'train': [...], 'val': [...]} with the reference's accessors (utils.py:43-99)."""

	def __init__(self, container: Dict[str, List[float]] = None):
		self.container = defaultdict(list)
		if container is not None:
			self.container.update(container)

	def __getitem__(self, item):
		return self.container[item]

	def __setitem__(self, key, value):
		self.container[key] = value

	def __contains__(self, item):
		return item in self.container

	def __iter__(self):
		return iter(self.container)

	def __len__(self):
		return len(self.container)

	def items(self):
		return self.container.items()

	def concat(self, other):
		for key, values in other.items():
			if isinstance(values, list):
				self.container[key].extend(values)
			else:
				self.container[key].append(values)

	def append(self, key, value):
		self.container[key].append(value)

	def min(self, key="val"):
		return min(self[key]) if key in self and len(self[key]) else np.inf

	def min_item(self, key="val"):
		if key in self:
			i = int(np.argmin(self[key]))
			return {k: v[i] for k, v in self.items()}

	def plot(self, save_path=None, show=False):
		try:
			import matplotlib
			matplotlib.use("Agg")
			import matplotlib.pyplot as plt
		except ImportError:
			return False
		fig, ax = plt.subplots(figsize=(12, 10))
		for name, values in self.items():
			ax.plot(values, label=name, linewidth=3)
		ax.set_xlabel("Epoch [-]", fontsize=16)
		ax.set_ylabel("Loss [-]", fontsize=16)
		ax.legend(fontsize=16)
		if save_path is not None:
			plt.savefig(save_path, dpi=300)
		if show:
			plt.show()
		plt.close(fig)
		return True
