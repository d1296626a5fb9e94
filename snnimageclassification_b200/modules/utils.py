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
	"""Deep merge of mapping ``u`` into ``d`` (in place, returned): nested mappings are merged key by key, any other
	value replaces what was there.  Used for the checkpoint index (same contract as the reference's helper,
	utils.py:28-40); written as an explicit work list rather than by recursion."""
	work = [(d, u)]
	while work:
		dst, src = work.pop()
		for key in src:
			new = src[key]
			if isinstance(new, collections.abc.Mapping):
				cur = dst.get(key)
				if not isinstance(cur, collections.abc.MutableMapping):
					cur = dst[key] = {}
				work.append((cur, new))
			else:
				dst[key] = new
	return d


class LossHistory:
	"""Per-phase loss curves, e.g. ``{"train": [...], "val": [...]}``.

	Interface of the reference's container (utils.py:43-99: ``container`` attribute, item access, ``items``,
	``concat``, ``append``, ``min``, ``min_item``, ``plot``) so that ``SNN.fit`` and user code keep working; a missing
	phase reads as an empty curve.
	"""

	def __init__(self, container: Dict[str, List[float]] = None):
		self.container = defaultdict(list)
		for phase, curve in (container or {}).items():
			self.container[phase] = list(curve)

	# -- mapping protocol -----------------------------------------------------------------------------------------------
	def __len__(self):
		return len(self.container)

	def __iter__(self):
		yield from self.container

	def __contains__(self, phase):
		return phase in self.container

	def __getitem__(self, phase):
		return self.container[phase]

	def __setitem__(self, phase, curve):
		self.container[phase] = curve

	def items(self):
		return self.container.items()

	# -- growing the curves ---------------------------------------------------------------------------------------------
	def append(self, phase, value):
		self.container[phase].append(value)

	def concat(self, other):
		"""Adds one epoch (``{"train": 0.3, "val": 0.4}``) or a whole history (lists) phase by phase."""
		for phase, values in other.items():
			curve = self.container[phase]
			curve.extend(values) if isinstance(values, list) else curve.append(values)

	# -- queries --------------------------------------------------------------------------------------------------------
	def min(self, key="val"):
		curve = self.container[key] if key in self.container else ()
		return min(curve) if len(curve) else np.inf

	def min_item(self, key="val"):
		"""Values of every phase at the epoch where ``key`` is smallest (None if ``key`` was never recorded)."""
		if key not in self.container:
			return None
		best = int(np.argmin(self.container[key]))
		return {phase: curve[best] for phase, curve in self.container.items()}

	def plot(self, save_path=None, show=False):
		"""Loss curves as a figure; returns False (and does nothing) when matplotlib is not installed."""
		try:
			import matplotlib
			matplotlib.use("Agg")
			from matplotlib import pyplot
		except ImportError:
			return False
		figure, axes = pyplot.subplots(figsize=(12, 10))
		for phase, curve in self.container.items():
			axes.plot(curve, linewidth=3, label=phase)
		axes.set(xlabel="Epoch [-]", ylabel="Loss [-]")
		axes.xaxis.label.set_size(16)
		axes.yaxis.label.set_size(16)
		axes.legend(fontsize=16)
		if save_path is not None:
			figure.savefig(save_path, dpi=300)
		if show:
			pyplot.show()
		pyplot.close(figure)
		return True
