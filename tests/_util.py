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


def elementwise_err(a, b, floor_frac=1e-3):
	"""max over elements of |a-b| / max(|b|, floor), floor = floor_frac * max|b|: an ELEMENTWISE relative error with an
	absolute floor, so that small-magnitude entries are checked too (``rel_err`` above is a max-norm bar: it only sees
	the entries near the largest magnitude)."""
	a = np.asarray(a, dtype=np.float64)
	b = np.asarray(b, dtype=np.float64)
	floor = max(np.abs(b).max() * floor_frac, 1e-30)
	return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


def unexplained_forks(Z_test, Z_ref, V_ref, thr_ref, tol=2e-5):
	"""Rasters of a chaotic recurrent net can only be compared up to the first spike that sits ON the threshold: a
	membrane potential within rounding distance of its threshold flips under ANY change of summation order (also
	between the reference on CPU and on GPU), after which that sample follows another trajectory.

	For every sample whose raster differs from the reference's, look at the FIRST differing step: every neuron that
	differs there must be such a near-tie in the reference's own trace, |V - thr| <= tol * max(|thr|, |V|).  Returns
	(number of forked samples, number of forked samples whose first difference is NOT a near-tie).  The second number
	must be 0; the first is reported and bounded separately."""
	Z_test, Z_ref = np.asarray(Z_test), np.asarray(Z_ref)
	B, T, H = Z_ref.shape
	thr_ref = np.broadcast_to(np.asarray(thr_ref, dtype=np.float32), Z_ref.shape)
	first = first_divergence(Z_test, Z_ref)
	forked = unexplained = 0
	for b in np.nonzero(first < T)[0]:
		t = first[b]
		forked += 1
		diff = Z_test[b, t] != Z_ref[b, t]
		v, th = V_ref[b, t][diff].astype(np.float64), thr_ref[b, t][diff].astype(np.float64)
		if np.any(np.abs(v - th) > tol * np.maximum(np.abs(th), np.abs(v))):
			unexplained += 1
	return forked, unexplained
