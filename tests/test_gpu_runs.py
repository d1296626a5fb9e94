"""Frame-dedup fast path: the encoder's run table against a numpy restatement, and the compact
kernels of snnk_forward / snnk_backward against the dense ones on the same inputs and weights."""
import numpy as np
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def npy(t):
	return t.detach().cpu().numpy()


def _images(B, N, seed, ink=0.19):
	g = torch.Generator().manual_seed(seed)
	k = torch.randint(1, 256, (B, N), generator=g).float() / 255.0
	return torch.where(torch.rand(B, N, generator=g) < ink, k, torch.zeros(()))


def _table_from_raster(x):
	B, T, _ = x.shape
	cap = max(128, (B * T // 4 + 127) // 128 * 128)
	row2c = np.zeros(B * T, dtype=np.int64)
	rep, ln = [], []
	for b in range(B):
		for t in range(T):
			if t == 0 or (x[b, t] != x[b, t - 1]).any():
				rep.append(b * T + t)
				ln.append(0)
			ln[-1] += 1
			row2c[b * T + t] = len(rep) - 1
	return cap, row2c, np.array(rep), np.array(ln)


@pytest.mark.parametrize("tau,periodic,T", [(0.02, True, 100), (0.02, False, 40), (20.0, True, 50), (20.0, False, 33)])
def test_run_table_matches_raster(tau, periodic, T):
	from snnimageclassification_b200 import ToSpikes
	from snnimageclassification_b200.modules.functional import get_runs
	enc = ToSpikes(T, tau=tau, use_periods=periodic)
	x = enc.encode_batch(_images(23, 196, 1).to(DEV))
	table = get_runs(x)
	assert table is not None
	tab = table.cpu().numpy()
	cap, row2c, rep, ln = _table_from_raster(npy(x))
	B = x.shape[0]
	assert tab[2] == cap and tab[0] == len(rep) and tab[1] == int(len(rep) <= cap) and tab[3] == 0
	assert np.array_equal(tab[4:4 + B * T], row2c)
	n = min(len(rep), cap)
	assert np.array_equal(tab[4 + B * T: 4 + B * T + n], rep[:n])
	m = n if tab[1] else n - 1       # the last stored run of an overfull table may be cut short
	assert np.array_equal(tab[4 + B * T + cap: 4 + B * T + cap + m], ln[:m])
	# the raster itself is unchanged by recording the runs
	plain = enc.encode_batch(_images(23, 196, 1).to(DEV), frame_runs=False)
	assert torch.equal(plain, x) and get_runs(plain) is None


def _net(N, H, T, seed=0):
	from snnimageclassification_b200 import LayerType, SNN, SpikeFuncType
	torch.manual_seed(seed)
	return SNN(N, 10, H, use_recurrent_connection=True, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True)


def _step(net, x, y):
	net.zero_grad()
	loss = net.batch_loss(x, y, torch.nn.NLLLoss())
	loss.backward()
	return float(loss), {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("B,T,N,H,rec", [(64, 100, 784, 128, True), (37, 50, 196, 64, True), (16, 100, 784, 128, False),
	(5, 7, 20, 32, True), (768, 16, 64, 128, True), (300, 130, 48, 128, True)])   # last two: MMA recurrence (expanded
	# compact rows); T > 128 (single-CTA run-table build)
def test_dedup_matches_dense(B, T, N, H, rec):
	"""Production encoder (tau = 0.02, periodic): <= 3 runs per sample -> the compact kernels run.  The projection of a
	run's first row is the same MMA sequence as in the dense kernel, so the forward pass is bit-identical; the weight
	gradients differ by summation order only."""
	from snnimageclassification_b200 import LayerType, SNN, SpikeFuncType, ToSpikes
	from snnimageclassification_b200.modules.functional import get_runs, mark_binary
	enc = ToSpikes(T, use_periods=True)
	x = enc.encode_batch(_images(B, N, 3).to(DEV))
	assert int(get_runs(x)[1]) == 1 and int(get_runs(x)[0]) <= 3 * B
	x_dense = mark_binary(x.clone())
	assert get_runs(x_dense) is None
	y = torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(4)).to(DEV)
	torch.manual_seed(0)
	net = SNN(N, 10, H, use_recurrent_connection=rec, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
		hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True)
	net.train()
	out_a, hid_a = net(x)
	out_b, hid_b = net(x_dense)
	assert torch.equal(out_a, out_b)
	for u, v in zip(hid_a["input"], hid_b["input"]):
		assert torch.equal(u, v)
	la, ga = _step(net, x, y)
	lb, gb = _step(net, x_dense, y)
	assert la == lb
	for k in gb:
		assert rel_err(npy(ga[k]), npy(gb[k])) <= 1e-4, k


def test_many_runs_take_the_dense_kernels():
	"""tau = 20 spreads the periods: more runs than the table holds -> ok = 0 -> the dense kernels run, bit-identically."""
	from snnimageclassification_b200 import ToSpikes
	from snnimageclassification_b200.modules.functional import get_runs, mark_binary
	B, T, N, H = 32, 100, 784, 128
	enc = ToSpikes(T, tau=20.0, use_periods=True)
	x = enc.encode_batch(_images(B, N, 5).to(DEV))
	tab = get_runs(x)
	assert int(tab[1]) == 0 and int(tab[0]) > int(tab[2])
	y = torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(6)).to(DEV)
	net = _net(N, H, T)
	net.train()
	la, ga = _step(net, x, y)
	lb, gb = _step(net, mark_binary(x.clone()), y)
	assert la == lb
	for k in gb:
		assert torch.equal(ga[k], gb[k]), k


def test_dedup_in_graphed_step_and_fp32_mode():
	"""The choice between the variants is made on the device, so one captured graph serves batches of either kind;
	with tensor_core=False the table is ignored (fp32 SIMT kernels, bit-exact path)."""
	from snnimageclassification_b200 import FusedAdam, LayerType, SNN, SpikeFuncType, ToSpikes
	B, T, N, H = 32, 100, 784, 128
	enc = ToSpikes(T, use_periods=True)
	imgs = [_images(B, N, 10 + i) for i in range(3)]
	y = torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(7))
	def run(graphs):
		torch.manual_seed(0)
		net = SNN(N, 10, H, use_recurrent_connection=True, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
			hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True, input_encoder=enc, cuda_graphs=graphs)
		opt = FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
		net.train()
		losses = [net._exec_batch(imgs[i % 3], y, torch.nn.NLLLoss(), opt) for i in range(5)]
		return losses, [p.detach().clone() for p in net.parameters()]
	l_eager, p_eager = run(False)
	l_graph, p_graph = run(True)
	assert np.allclose(l_eager, l_graph, rtol=1e-5)
	for a, b in zip(p_eager, p_graph):
		assert rel_err(npy(b), npy(a)) <= 1e-4
	torch.manual_seed(0)
	exact = SNN(N, 10, H, use_recurrent_connection=True, int_time_steps=T, hidden_layer_type=LayerType.ALIF, device=DEV,
		learn_beta=True, tensor_core=False)
	x = enc.encode_batch(imgs[0].to(DEV))
	from snnimageclassification_b200.modules.functional import mark_binary
	a, _ = exact(x)
	b, _ = exact(mark_binary(x.clone()))
	assert torch.equal(a, b)


def test_lazy_raster_training_matches_dense_raster():
	"""SNN's own intermediate raster is written lazily (only the rows the dedup kernels read); training through
	_exec_batch must not notice -- production encoder (dedup variant runs) and tau = 20 (table not ok: every row is
	written and the dense kernels run), plus a wide first layer (variant not eligible: never lazy)."""
	import os
	from snnimageclassification_b200 import FusedAdam, LayerType, SNN, SpikeFuncType, ToSpikes
	B, T, N = 32, 100, 784
	imgs = [_images(B, N, 20 + i) for i in range(3)]
	y = torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(8))
	for tau, H in ((0.02, 128), (20.0, 128), (0.02, 256)):
		enc = ToSpikes(T, tau=tau, use_periods=True)
		res = {}
		for lazy in ("1", "0"):
			os.environ["SNNK_LAZY_RASTER"] = lazy
			torch.manual_seed(0)
			net = SNN(N, 10, H, use_recurrent_connection=True, int_time_steps=T, spike_func=SpikeFuncType.FastSigmoid,
				hidden_layer_type=LayerType.ALIF, device=DEV, learn_beta=True, input_encoder=enc)
			opt = FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
			net.train()
			losses = [net._exec_batch(imgs[i % 3], y, torch.nn.NLLLoss(), opt) for i in range(4)]
			res[lazy] = (losses, [p.detach().clone() for p in net.parameters()])
		os.environ.pop("SNNK_LAZY_RASTER")
		assert res["1"][0] == res["0"][0], (tau, H)
		for a, b in zip(res["1"][1], res["0"][1]):
			assert torch.equal(a, b), (tau, H)


@pytest.mark.parametrize("tau,periodic,N", [(0.02, True, 784), (20.0, True, 100), (20.0, False, 33)])
def test_bit_packed_raster_format(tau, periodic, N):
	"""SNNK_BITS: bit l of word w = pixel 32 w + l; unpacking gives the dense raster back; SNN takes the
	packed raster directly and computes what it computes on the dense one."""
	from snnimageclassification_b200 import LayerType, SNN, ToSpikes, unpack_raster
	T, B = 24, 9
	enc = ToSpikes(T, tau=tau, use_periods=periodic)
	img = _images(B, N, 31).to(DEV)
	dense = enc.encode_batch(img, frame_runs=False)
	bits = enc.encode_batch_bits(img)
	assert bits.dtype == torch.int32 and tuple(bits.shape) == (B, T, (N + 31) // 32)
	d, w = npy(dense).astype(np.uint32), npy(bits).view(np.uint32)
	ref = np.zeros_like(w)
	for c in range(N):
		ref[..., c // 32] |= d[..., c] << np.uint32(c % 32)
	assert np.array_equal(w, ref)
	assert torch.equal(unpack_raster(bits, N), dense)
	torch.manual_seed(0)
	net = SNN(N, 10, 32, int_time_steps=T, hidden_layer_type=LayerType.LIF, device=DEV, tensor_core=False)
	a, _ = net(dense)
	b, _ = net(bits.cpu())          # e.g. a packed raster coming from a host-side data loader
	assert torch.equal(a, b)


def test_full_size_determinism_and_row_independence():
	"""BASELINE configs[1] geometry (B = 256, T = 100, 784-128-10, production encoder, tensor-core + dedup kernels):
	two identical training steps give bit-identical loss and gradients (every reduction has a fixed order), and
	permuting the batch rows permutes the traces bit-exactly (rows never interact before the loss mean)."""
	from snnimageclassification_b200 import ToSpikes
	B, T, N, H = 256, 100, 784, 128
	enc = ToSpikes(T, use_periods=True)
	img = _images(B, N, 41)
	y = torch.randint(0, 10, (B,), generator=torch.Generator().manual_seed(42)).to(DEV)
	x = enc.encode_batch(img.to(DEV))
	net = _net(N, H, T)
	net.train()
	l1, g1 = _step(net, x, y)
	l2, g2 = _step(net, x, y)
	assert l1 == l2
	for k in g1:
		assert torch.equal(g1[k], g2[k]), k
	perm = torch.randperm(B, generator=torch.Generator().manual_seed(43))
	xp = enc.encode_batch(img[perm].to(DEV))
	out, hid = net(x)
	outp, hidp = net(xp)
	assert torch.equal(outp, out[perm.to(DEV)])
	for u, v in zip(hidp["input"], hid["input"]):
		assert torch.equal(u, v[perm.to(DEV)])


def test_tiled_compact_rows_behind_the_table(monkeypatch):
	"""encode_batch leaves the first row of every run behind the run table, tiled for the compact projection
	(SNNK_F_RUNS_TILED): the step then skips its gather.  Same results, bit for bit, as with the gather."""
	from snnimageclassification_b200 import ToSpikes
	from snnimageclassification_b200.modules import functional as F_
	B, T, N, H, O = 48, 40, 784, 128, 10
	img = _images(B, N, 3).to(DEV)
	enc = ToSpikes(T, use_periods=True)
	x_t = enc.encode_batch(img)
	monkeypatch.setenv("SNNK_RUNS_TILED", "0")
	x_g = enc.encode_batch(img)
	monkeypatch.delenv("SNNK_RUNS_TILED")
	assert F_.runs_tiled(x_t) and not F_.runs_tiled(x_g) and F_.get_runs(x_g) is not None
	assert torch.equal(x_t, x_g)
	ta, tb = F_.get_runs(x_t).cpu().numpy(), F_.get_runs(x_g).cpu().numpy()
	n_rows, cap = int(tb[0]), int(tb[2])
	assert np.array_equal(ta[:4 + B * T + n_rows], tb[:4 + B * T + n_rows])                     # header, row -> run, first rows
	assert np.array_equal(ta[4 + B * T + cap: 4 + B * T + cap + n_rows], tb[4 + B * T + cap: 4 + B * T + cap + n_rows])   # run lengths
	g = torch.Generator().manual_seed(0)
	W_in = (torch.randn(N, H, generator=g) * 0.03).to(DEV)
	W_rec = (torch.randn(H, H, generator=g) * 0.03).to(DEV)
	mask = (1 - torch.eye(H)).to(DEV)
	W_out, b_out = torch.randn(H, O, generator=g).to(DEV), torch.zeros(O, device=DEV)
	beta = torch.tensor([1.6], device=DEV)
	c = F_.LayerConsts(1, 0, True, 0.95, 0.995, 0.03, 0.3, 0.9, tensor_core=True)
	f_t = F_.run_forward(c, x_t, W_in, W_rec, mask, beta, W_out, b_out)
	f_g = F_.run_forward(c, x_g, W_in, W_rec, mask, beta, W_out, b_out)
	assert float(f_t["Z"].mean()) > 0.0
	for k in ("V", "a", "Z", "y", "logits"):
		assert torch.equal(f_t[k], f_g[k]), k
