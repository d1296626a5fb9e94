"""Shared helpers for the test-suite: fixture loading and parity metrics."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
	return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack_bits(bits, shape):
	n = int(np.prod(shape))
	return np.unpackbits(bits)[:n].reshape(shape)


def dynamics_case(z, name):
	"""Returns the dict stored for one dynamics case."""
	pre = name + "/"
	return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def rel_err(a, b):
	"""max |a-b| / max(|b|, tiny): the relative measure used for gradients and losses."""
	a = np.asarray(a, dtype=np.float64)
	b = np.asarray(b, dtype=np.float64)
	den = max(np.abs(b).max(), 1e-30)
	return float(np.abs(a - b).max() / den)


def first_divergence(Za, Zb):
	"""Per sample, the first time step at which two rasters (B,T,H) differ (T if never)."""
	diff = (np.asarray(Za) != np.asarray(Zb)).any(axis=2)
	T = diff.shape[1]
	return np.where(diff.any(axis=1), diff.argmax(axis=1), T)
