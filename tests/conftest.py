import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
	"""GPU tests fail loudly (not skip) on a box without CUDA only when explicitly selected with -m gpu."""
	import torch
	if torch.cuda.is_available():
		return
	selected = config.getoption("-m") or ""
	if "gpu" in selected and "not gpu" not in selected:
		return
	skip = pytest.mark.skip(reason="no CUDA device in this container")
	for item in items:
		if "gpu" in item.keywords:
			item.add_marker(skip)
