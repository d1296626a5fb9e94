"""CPU oracle for the spiking hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this package.  The product package never does.

Two restatements live here:

* ``snn_oracle.c`` (loaded through ctypes below): scalar fp32 C, fixed
  summation orders, used as the bit-level checker of the CUDA kernels.
* ``torch_port.py``: the same algorithm written with PyTorch CPU ops and
  autograd, i.e. what the reference actually executes; used as the second
  checker and as the timed CPU baseline.

Parity status: pinned against the reference (tests/golden/, see
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsnn_oracle.so")
_lib = None


def build(force: bool = False) -> str:
	"""Compile ``snn_oracle.c`` with the committed Makefile."""
	src = os.path.join(_HERE, "snn_oracle.c")
	stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
	if force or stale:
		subprocess.run(["make", "-C", _HERE, "-B", "CC=gcc"], check=True, capture_output=True)
	return _LIB_PATH


class _Cfg(ctypes.Structure):
	_fields_ = [
		("B", ctypes.c_int32), ("T", ctypes.c_int32), ("N", ctypes.c_int32),
		("H", ctypes.c_int32), ("O", ctypes.c_int32),
		("layer_type", ctypes.c_int32), ("surrogate", ctypes.c_int32), ("recurrent", ctypes.c_int32),
		("alpha", ctypes.c_float), ("rho", ctypes.c_float), ("theta", ctypes.c_float),
		("gamma", ctypes.c_float), ("kappa", ctypes.c_float), ("beta", ctypes.c_float),
		("dt", ctypes.c_float), ("iz_C", ctypes.c_float), ("iz_vr", ctypes.c_float), ("iz_vth", ctypes.c_float),
		("iz_k", ctypes.c_float), ("iz_a", ctypes.c_float), ("iz_b", ctypes.c_float), ("iz_c", ctypes.c_float),
		("iz_d", ctypes.c_float), ("iz_vpeak", ctypes.c_float),
	]


@dataclass
class OracleCfg:
	B: int
	T: int
	N: int
	H: int
	O: int
	layer_type: int = 1  # 0 LIF, 1 ALIF, 2 Izhikevich
	surrogate: int = 0  # 0 FastSigmoid, 1 Phi
	recurrent: int = 1
	alpha: float = 0.0
	rho: float = 0.0
	theta: float = 1.0
	gamma: float = 1.0
	kappa: float = 0.0
	beta: float = 0.0
	# Izhikevich constants (defaults of spiking_layers.py:287-296)
	dt: float = 1e-3
	iz_C: float = 100.0
	iz_vr: float = -60.0
	iz_vth: float = -40.0
	iz_k: float = 0.7
	iz_a: float = 0.03
	iz_b: float = -2.0
	iz_c: float = -50.0
	iz_d: float = 100.0
	iz_vpeak: float = 35.0

	def c(self) -> _Cfg:
		return _Cfg(
			self.B, self.T, self.N, self.H, self.O, self.layer_type, self.surrogate, int(self.recurrent),
			self.alpha, self.rho, self.theta, self.gamma, self.kappa, self.beta,
			self.dt, self.iz_C, self.iz_vr, self.iz_vth, self.iz_k, self.iz_a, self.iz_b, self.iz_c, self.iz_d,
			self.iz_vpeak,
		)


def lib():
	global _lib
	if _lib is None:
		build()
		_lib = ctypes.CDLL(_LIB_PATH)
	return _lib


def _p(a: Optional[np.ndarray], ctype=ctypes.c_float):
	if a is None:
		return ctypes.cast(None, ctypes.POINTER(ctype))
	assert a.flags["C_CONTIGUOUS"]
	return a.ctypes.data_as(ctypes.POINTER(ctype))


def _f32(a) -> Optional[np.ndarray]:
	if a is None:
		return None
	return np.ascontiguousarray(a, dtype=np.float32)


def periods(x: np.ndarray, t_max: float, tau: float, thr: float, eps: float) -> np.ndarray:
	"""datasets.py:42-54 for a float32 or float64 array of any shape."""
	x = np.ascontiguousarray(x)
	out = np.empty(x.shape, dtype=np.int64)
	if x.dtype == np.float64:
		lib().snn_oracle_periods_f64(
			_p(x, ctypes.c_double), ctypes.c_int64(x.size), ctypes.c_double(t_max), ctypes.c_double(tau),
			ctypes.c_double(thr), ctypes.c_double(eps), _p(out, ctypes.c_int64))
	elif x.dtype == np.float32:
		lib().snn_oracle_periods_f32(
			_p(x, ctypes.c_float), ctypes.c_int64(x.size), ctypes.c_double(t_max), ctypes.c_double(tau),
			ctypes.c_double(thr), ctypes.c_double(eps), _p(out, ctypes.c_int64))
	else:
		raise TypeError(x.dtype)
	return out


def raster(per: np.ndarray, n_steps: int, periodic: bool) -> np.ndarray:
	"""datasets.py:72-86.  per: (n_items, n_pix) int64 -> (n_items, n_steps, n_pix) uint8."""
	per = np.ascontiguousarray(per, dtype=np.int64)
	n_items, n_pix = per.shape
	out = np.empty((n_items, n_steps, n_pix), dtype=np.uint8)
	lib().snn_oracle_raster(
		_p(per, ctypes.c_int64), ctypes.c_int64(n_items), ctypes.c_int64(n_pix), ctypes.c_int32(n_steps),
		ctypes.c_int32(int(periodic)), _p(out, ctypes.c_uint8))
	return out


def encode(x: np.ndarray, n_steps: int, t_max=None, tau=20.0 * 1e-3, thr=0.2, periodic=False, eps=1e-7) -> np.ndarray:
	"""ToSpikes.__call__ (datasets.py:93-97) for a batch: x (n_items, n_pix) -> (n_items, n_steps, n_pix) uint8."""
	x = np.ascontiguousarray(x)
	if x.ndim == 1:
		return encode(x[None], n_steps, t_max, tau, thr, periodic, eps)[0]
	t_max = n_steps if t_max is None else t_max
	return raster(periods(x, t_max, tau, thr, eps), n_steps, periodic)


def forward(cfg: OracleCfg, x, W_in, W_rec, rec_mask, W_out, b_out, V0=None, a0=None, Z0=None):
	"""Returns dict(I_in, V, a, Z, y) of float32 arrays."""
	B, T, H, O = cfg.B, cfg.T, cfg.H, cfg.O
	x, W_in, W_rec, rec_mask, W_out, b_out = map(_f32, (x, W_in, W_rec, rec_mask, W_out, b_out))
	V0, a0, Z0 = map(_f32, (V0, a0, Z0))
	if cfg.layer_type == 2 and V0 is None:     # IzhikevichLayer.create_empty_state: V starts at v_rest (spiking_layers.py:309)
		V0 = np.full((B, H), cfg.iz_vr, np.float32)
	out = {k: np.zeros((B, T, H), np.float32) for k in ("I_in", "V", "a", "Z")}
	out["y"] = np.zeros((B, T, O), np.float32)
	c = cfg.c()
	rc = lib().snn_oracle_forward(
		ctypes.byref(c), _p(x), _p(W_in), _p(W_rec), _p(rec_mask), _p(W_out), _p(b_out),
		_p(V0), _p(a0), _p(Z0), _p(out["I_in"]), _p(out["V"]), _p(out["a"]), _p(out["Z"]), _p(out["y"]))
	if rc != 0:
		raise RuntimeError(f"snn_oracle_forward failed: {rc}")
	return out


def head(y: np.ndarray, labels: Optional[np.ndarray]):
	"""Returns dict(logits, tstar, logp, loss, g_y)."""
	y = _f32(y)
	B, T, O = y.shape
	lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.int64)
	out = dict(
		logits=np.zeros((B, O), np.float32), tstar=np.zeros((B, O), np.int32), logp=np.zeros((B, O), np.float32),
		loss=np.zeros((1,), np.float32), g_y=np.zeros((B, T, O), np.float32),
	)
	rc = lib().snn_oracle_head(
		ctypes.c_int(B), ctypes.c_int(T), ctypes.c_int(O), _p(y), _p(lab, ctypes.c_int64), _p(out["logits"]),
		_p(out["tstar"], ctypes.c_int32), _p(out["logp"]), _p(out["loss"]), _p(out["g_y"]))
	if rc != 0:
		raise RuntimeError(f"snn_oracle_head failed: {rc}")
	out["loss"] = float(out["loss"][0])
	return out


def backward(cfg: OracleCfg, x, W_rec, rec_mask, W_out, V, a, Z, g_y, Z0=None, g_Vs=None, g_Zs=None):
	"""Returns dict(gI, dW_in, dW_rec, dW_out, db)."""
	B, T, N, H, O = cfg.B, cfg.T, cfg.N, cfg.H, cfg.O
	x, W_rec, rec_mask, W_out, V, a, Z, g_y, Z0, g_Vs, g_Zs = map(
		_f32, (x, W_rec, rec_mask, W_out, V, a, Z, g_y, Z0, g_Vs, g_Zs))
	out = dict(
		gI=np.zeros((B, T, H), np.float32), dW_in=np.zeros((N, H), np.float32),
		dW_rec=np.zeros((H, H), np.float32) if cfg.recurrent else None,
		dW_out=np.zeros((H, O), np.float32), db=np.zeros((O,), np.float32),
	)
	c = cfg.c()
	rc = lib().snn_oracle_backward(
		ctypes.byref(c), _p(x), _p(W_rec), _p(rec_mask), _p(W_out), _p(Z0), _p(V), _p(a), _p(Z), _p(g_y),
		_p(g_Vs), _p(g_Zs), _p(out["gI"]), _p(out["dW_in"]), _p(out["dW_rec"]), _p(out["dW_out"]), _p(out["db"]))
	if rc != 0:
		raise RuntimeError(f"snn_oracle_backward failed: {rc}")
	return out
