"""Parity of the CUDA path (through the reference-facing modules and the C ABI)
against the reference's own outputs (tests/golden) and the CPU oracle.

Bars (BASELINE.json north_star): volume rel-L2 <= 1e-5, gradient rel-L2 <= 1e-4.
"""
import numpy as np
import pytest
import torch

from oracle import lct_oracle as O
from tests._golden import Case, case_names, case_shape, numpy_lct_fixture

pytestmark = pytest.mark.gpu

TOL_Y = 1e-5       # relative L2, fp32 volume
TOL_G = 1e-4       # relative L2, gradient


def _layer(N, M, bin_len, D, method="lct", material="diffuse"):
    import hiddenpose_b200 as hp
    layer = hp.lct(spatial=N, crop=M, bin_len=bin_len, wall_size=2.0, method=method, material=material)
    layer.todev("cuda:0", D)
    return layer


def test_native_library_is_loaded():
    from hiddenpose_b200 import _native
    lib = _native.load()
    assert lib.lct_abi_version() == 3
    with open("/proc/self/maps") as f:
        assert "libhiddenpose_lct.so" in f.read()


@pytest.mark.parametrize("name", case_names())
def test_golden_forward_backward(name, parity_log):
    """CUDA forward + backward vs the outputs of the reference itself (tflct.py:94-179 run by
    tests/golden/make_golden.py) -- every BASELINE.json shape is among the cases: 256x64x64 (configs 1-2),
    512x128x128 (config 3), 128x128x128 (config 4), 512x256x256 (config 5)."""
    c = Case(name)
    layer = _layer(c.N, c.M, c.bin_len, c.D, c.method, c.material)
    x = torch.from_numpy(c.x).cuda().requires_grad_(True)
    y = layer(x, c.tbes, c.tens)
    assert y.shape == (c.B, c.D, c.M, c.N, c.N)
    y.backward(torch.from_numpy(c.g).cuda())
    ey, eg = c.y_err(y.detach().cpu().numpy()), c.gx_err(x.grad.cpu().numpy())
    parity_log(M=c.M, N=c.N, volume_rel_l2=float(ey), grad_rel_l2=float(eg), against="reference golden")
    assert ey <= TOL_Y
    assert eg <= TOL_G


@pytest.mark.parametrize("M,N,B,D", [(32, 8, 3, 2), (64, 16, 2, 3), (128, 32, 2, 1), (64, 64, 1, 2), (32, 128, 1, 1)])
def test_vs_oracle_ragged_windows(M, N, B, D):
    """Per-sample windows (different tbe per sample), D > 1, against the CPU oracle."""
    bl = 0.01 * 512 / M
    layer = _layer(N, M, bl, D)
    orc = O.LctOracle(N, M, bl)
    rs = np.random.RandomState(M + N)
    tin = M - 9
    tbes = [int(v) for v in rs.randint(0, 10, size=B)]
    tens = [t + tin for t in tbes]
    x = torch.from_numpy(rs.rand(B, D, tin, N, N).astype(np.float32))
    g = torch.from_numpy(rs.randn(B, D, M, N, N).astype(np.float32))
    xd = x.cuda().requires_grad_(True)
    y = layer(xd, tbes, tens)
    y.backward(g.cuda())
    xo = x.clone().requires_grad_(True)
    yo = orc.forward(xo, tbes, tens)
    yo.backward(g)
    assert O.rel_l2(y.detach().cpu(), yo.detach()) <= TOL_Y
    assert O.rel_l2(xd.grad.cpu(), xo.grad) <= TOL_G


def test_chunked_workspace_matches_full():
    """A workspace that only fits 1 channel at a time gives the same bits as one batch."""
    M, N, B, D = 64, 16, 3, 2
    layer = _layer(N, M, 0.08, D)
    x = torch.rand(B, D, M, N, N, device="cuda")
    y_full = layer(x, [0] * B, [M] * B)
    plan = layer._plan
    saved = plan.workspace_limit_bytes
    try:
        plan.workspace_limit_bytes = 1
        y_chunk = layer(x, [0] * B, [M] * B)
    finally:
        plan.workspace_limit_bytes = saved
    assert torch.equal(y_full, y_chunk)


def test_zero_and_linearity():
    M, N = 64, 16
    layer = _layer(N, M, 0.08, 1)
    z = layer(torch.zeros(1, 1, M, N, N, device="cuda"), [0], [M])
    assert float(z.abs().max()) == 0.0
    a, b = torch.rand(1, 1, M, N, N, device="cuda"), torch.rand(1, 1, M, N, N, device="cuda")
    ya, yb, yab = layer(a, [0], [M]), layer(b, [0], [M]), layer(2.0 * a - 3.0 * b, [0], [M])
    assert O.rel_l2((2.0 * ya - 3.0 * yb).cpu(), yab.cpu()) <= 1e-5


def test_impulse_matches_oracle_column():
    """A single impulse reproduces the oracle's operator column (window offset included)."""
    M, N = 32, 8
    layer = _layer(N, M, 0.16, 1)
    orc = O.LctOracle(N, M, 0.16)
    x = torch.zeros(1, 1, 20, N, N)
    x[0, 0, 13, 5, 2] = 1.0
    y = layer(x.cuda(), [6], [26]).cpu()
    yo = orc.forward(x, [6], [26])
    assert O.rel_l2(y, yo) <= TOL_Y


@pytest.mark.parametrize("M,N,C", [(256, 64, 8), (128, 128, 4), (512, 128, 2), (512, 256, 1)])
def test_adjoint_identity_full_size(M, N, C):
    """<A x, g> == <x, A^T g> at the BASELINE.json shapes (size-independent property):
    ties the hand-written backward chain to the forward chain."""
    layer = _layer(N, M, 0.01 * 512 / M, 1)
    gen = torch.Generator(device="cuda").manual_seed(410)
    x = torch.rand(C, 1, M, N, N, device="cuda", generator=gen)
    g = torch.randn(C, 1, M, N, N, device="cuda", generator=gen)
    y = layer(x, [0] * C, [M] * C)
    gx = layer._plan.backward(g, [0] * C, [M] * C, M)
    lhs = float((y.double() * g.double()).sum())
    rhs = float((x.double() * gx.double()).sum())
    scale = float(y.double().norm() * g.double().norm())
    assert abs(lhs - rhs) <= 2e-6 * scale


@pytest.mark.parametrize("M,N,C", [(256, 64, 2), (512, 128, 1)])
def test_full_size_vs_oracle_on_gpu(M, N, C):
    """Whole-volume comparison at BASELINE shapes against the oracle's op sequence run with
    torch's own CUDA FFT/GEMM (test-only use of cuFFT; the product never calls it)."""
    bl = 0.01 * 512 / M
    layer = _layer(N, M, bl, 1)
    orc = O.LctOracle(N, M, bl)
    gen = torch.Generator(device="cuda").manual_seed(411)
    x = torch.rand(C, 1, M, N, N, device="cuda", generator=gen)
    y = layer(x, [0] * C, [M] * C)
    yo = torch.cat([orc.forward(x[i:i + 1], [0], [M]) for i in range(C)])
    assert O.rel_l2(y.cpu(), yo.cpu()) <= TOL_Y


def test_feature_propagation_dropin_and_broadcast_lists():
    """FeaturePropagation API as NlosPose.py:25-32,53 uses it: int device, 3-entry lists, B > 3."""
    import hiddenpose_b200 as hp
    M, N, B = 64, 16, 5
    fp = hp.FeaturePropagation(time_size=M, image_size=N, wall_size=2.0, bin_len=0.08, dnum=1, dev=0)
    assert dict(fp.state_dict()) == {} and list(fp.parameters()) == []
    x = torch.rand(B, 1, M, N, N, device="cuda")
    y = fp(x, [0, 0, 0], [M, M, M])
    orc = O.LctOracle(N, M, 0.08)
    yo = orc.forward(x.cpu(), [0] * B, [M] * B)
    assert O.rel_l2(y.cpu(), yo) <= TOL_Y
    f = hp.normalize_feature(y)
    assert float(f.min()) == 0.0 and abs(float(f.max()) - 10.0) < 1e-4


def test_host_buffer_entry_point():
    """lct_forward_host (the C-ABI call with host buffers) equals the device-buffer path."""
    M, N, B = 64, 16, 2
    layer = _layer(N, M, 0.08, 1)
    x = torch.rand(B, 1, M, N, N).pin_memory()
    y_host = layer._plan.forward_host(x, [0] * B, [M] * B)
    y_dev = layer(x.cuda(), [0] * B, [M] * B)
    assert torch.equal(y_host, y_dev.cpu())


def test_error_behaviour():
    M, N = 32, 8
    layer = _layer(N, M, 0.16, 1)
    x = torch.zeros(1, 1, M, N, N, device="cuda")
    with pytest.raises(AssertionError):
        layer(x, [-1], [M - 1])
    with pytest.raises(AssertionError):
        layer(x, [1], [M + 1])
    with pytest.raises(AssertionError):
        layer(torch.zeros(1, 1, M, N, N + 1, device="cuda"), [0], [M])
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 2, M, N, N, device="cuda"), [0], [M])       # D != dnum
    with pytest.raises(RuntimeError):
        layer(x.cpu(), [0], [M])                                          # no CPU fallback
    with pytest.raises(IndexError):
        layer(torch.zeros(3, 1, M - 1, N, N, device="cuda"), [0, 1], [M - 1, M])


def test_streamer_matches_batch_by_batch():
    """LctStreamer (overlapped upload / transform / download) returns exactly what the layer returns."""
    import hiddenpose_b200 as hp
    M, N, B = 64, 16, 2
    layer = _layer(N, M, 0.08, 1)
    xs = [torch.rand(B, 1, M, N, N).pin_memory() for _ in range(5)]
    ys = [torch.empty(B, 1, M, N, N).pin_memory() for _ in range(5)]
    hp.LctStreamer(layer, [0] * B, [M] * B, depth=2).run(xs, ys)
    for xh, yh in zip(xs, ys):
        assert torch.equal(yh, layer(xh.cuda(), [0] * B, [M] * B).cpu())


def _ref_normalize_feature(x):
    """feature_propagation.py:273-286, the reference's own op sequence (torch)."""
    b, c, d, h, w = x.shape
    k = x.reshape(b, c, -1)
    z = k - k.min(2, keepdim=True)[0]
    n = z / (z.max(2, keepdim=True)[0] + 1e-15)
    n = n * 10.
    return n.view(b, c, d, h, w)


@pytest.mark.parametrize("handoff", ["implicit", "explicit", "none"])
def test_normalize_feature_forward_backward(handoff):
    """normalize_feature on the library (row f1): bit-identical forward, gradient vs autograd of the
    reference expression; with min/max handed over by the LCT's last kernel (implicitly, per tensor object,
    or explicitly as a returned tuple) and with the stand-alone reduction pass."""
    import hiddenpose_b200 as hp
    from hiddenpose_b200.lct_function import recall_minmax
    M, N, B = 64, 16, 3
    fp = hp.FeaturePropagation(time_size=M, image_size=N, bin_len=0.08, dnum=1, dev=0)
    x = torch.rand(B, 1, M, N, N, device="cuda", requires_grad=True)
    if handoff == "explicit":
        y, keys = fp.method.forward_with_minmax(x, [0] * 3, [M] * 3)
        assert recall_minmax(y) is None                  # nothing implicit on this path
        out = hp.normalize_feature(y, minmax=keys)
    else:
        y = fp(x, [0] * 3, [M] * 3)
        assert recall_minmax(y) is not None
        if handoff == "none":
            y = y.clone()                                # a different tensor object: stand-alone reduction pass
            assert recall_minmax(y) is None
        out = hp.normalize_feature(y)
    ref_in = y.detach().clone().requires_grad_(True)
    ref = _ref_normalize_feature(ref_in)
    assert torch.equal(out, ref)
    g = torch.randn_like(out)
    (gy,) = torch.autograd.grad(out, y, g)
    ref.backward(g)
    assert O.rel_l2(gy.cpu(), ref_in.grad.cpu()) <= 1e-5
    assert float(y.detach().reshape(B, -1).min(1)[0].min()) < 0          # negatives do reach the min (reference quirk C8)
    assert torch.equal(fp.forward_normalized(x.detach(), [0] * 3, [M] * 3), out.detach())


def test_minmax_handoff_misses_after_any_change():
    """The remembered keys belong to one tensor object in one state: an in-place write, a view or a copy must miss
    (and then agree with torch), and a dead tensor's entry must not survive for whatever reuses its address."""
    import gc
    import hiddenpose_b200 as hp
    from hiddenpose_b200 import lct_function as LF
    M, N = 64, 16
    fp = hp.FeaturePropagation(time_size=M, image_size=N, bin_len=0.08, dnum=1, dev=0)
    x = torch.rand(2, 1, M, N, N, device="cuda")
    with torch.no_grad():
        y = fp(x, [0] * 2, [M] * 2)
        assert LF.recall_minmax(y) is not None
        assert LF.recall_minmax(y[:]) is None and LF.recall_minmax(y.view(2, 1, M, N, N)) is None
        y.mul_(-3.0)                                         # min and max swap
        assert LF.recall_minmax(y) is None
        assert torch.equal(hp.normalize_feature(y), _ref_normalize_feature(y))
        n_before = len(LF._minmax_registry)
        del y
        gc.collect()
        assert len(LF._minmax_registry) == n_before - 1


@pytest.mark.parametrize("bad", [float("nan"), float("inf"), float("-inf")])
def test_normalize_feature_propagates_nan_and_inf_like_torch(bad):
    """A diverged volume must poison the normalised output as torch.min / torch.max do in the reference
    (feature_propagation.py:276-282), through the stand-alone reduction and through the LCT's fused one."""
    import hiddenpose_b200 as hp
    M, N = 64, 16
    v = torch.randn(2, 1, M, N, N, device="cuda")
    v[1, 0, 17, 3, 5] = bad
    got, want = hp.normalize_feature(v), _ref_normalize_feature(v)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(torch.nan_to_num(got, nan=7.0), torch.nan_to_num(want, nan=7.0))
    # through the layer: a non-finite input sample spreads over its whole volume (it is one 3-D convolution)
    fp = hp.FeaturePropagation(time_size=M, image_size=N, bin_len=0.08, dnum=1, dev=0)
    x = torch.rand(2, 1, M, N, N, device="cuda")
    x[1, 0, 20, 4, 4] = bad
    with torch.no_grad():
        y = fp(x, [0] * 2, [M] * 2)
        got, want = hp.normalize_feature(y), _ref_normalize_feature(y.clone())
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert bool(torch.isnan(got[1]).all()) and not bool(torch.isnan(got[0]).any())
    assert torch.equal(got[0], want[0])


def test_layer_survives_pickle_and_deepcopy(tmp_path):
    """train.py:223 ends with torch.save(model, path) and EMA helpers deep-copy the model: the native plan is dropped
    from the state and rebuilt on first use; the copy computes the same bits."""
    import copy
    import pickle
    import hiddenpose_b200 as hp
    M, N = 64, 16
    fp = hp.FeaturePropagation(time_size=M, image_size=N, bin_len=0.08, dnum=1, dev="cuda:0")
    x = torch.rand(2, 1, M, N, N, device="cuda")
    want = fp(x, [0, 0], [M, M])
    clones = [copy.deepcopy(fp), pickle.loads(pickle.dumps(fp))]
    path = tmp_path / "model.pth"
    torch.save(fp, path)
    clones.append(torch.load(path, weights_only=False))
    for clone in clones:
        assert clone.method._plan is None                   # not carried over ...
        assert torch.equal(clone(x, [0, 0], [M, M]), want)  # ... rebuilt lazily on the recorded device
        assert clone.method._plan is not None and clone.method._plan is not fp.method._plan
        assert dict(clone.state_dict()) == {}
    assert torch.equal(fp(x, [0, 0], [M, M]), want)          # the original is untouched


def test_ragged_windows_replay_from_a_cuda_graph():
    """Per-sample windows are legal under stream capture: the window table travels as kernel parameters, not as a
    copy from a host array that is gone by replay time."""
    M, N, B = 64, 16, 3
    layer = _layer(N, M, 0.08, 1)
    tin = M - 8
    tbes = [0, 5, 8]
    tens = [t + tin for t in tbes]
    static_x = torch.rand(B, 1, tin, N, N, device="cuda")
    with torch.no_grad():
        layer(static_x, tbes, tens)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_y = layer(static_x, list(tbes), list(tens))    # temporaries: dead by replay
        for _ in range(2):
            fresh = torch.rand_like(static_x)
            static_x.copy_(fresh)
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(static_y, layer(fresh, tbes, tens))


def test_two_threads_share_one_plan_on_two_streams():
    """include/hiddenpose_lct.h: one plan may be driven by several threads on several streams at once (distinct
    workspaces).  Each caller stream gets its own internal side streams; results equal the serial ones."""
    import threading
    M, N, B = 64, 32, 4
    layer = _layer(N, M, 0.08, 1)
    xs = [torch.rand(B, 1, M, N, N, device="cuda") for _ in range(2)]
    with torch.no_grad():
        want = [layer(x, [0] * B, [M] * B).clone() for x in xs]
    torch.cuda.synchronize()
    errors, start = [], threading.Barrier(2)

    def worker(i):
        try:
            stream = torch.cuda.Stream()
            start.wait()
            with torch.no_grad(), torch.cuda.stream(stream):
                for _ in range(200):
                    y = layer(xs[i], [0] * B, [M] * B)
                stream.synchronize()
                if not torch.equal(y, want[i]):
                    errors.append(f"thread {i}: result differs")
        except Exception as exc:              # pragma: no cover
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("N,M,B", [(64, 64, 6), (128, 32, 3)])
def test_dependent_launch_keeps_stream_order(monkeypatch, N, M, B):
    """The kernels of a chain are launched programmatically dependent (LCT_PDL, default on): each may start while its
    predecessor drains and orders itself with griddepcontrol.wait.  Back-to-back calls that reuse one workspace (the
    next call's first kernel overwrites the spectrum the previous call's last kernel reads), forward and backward
    mixed, must give the bits of a plan that launches in plain stream order."""
    monkeypatch.setenv("LCT_PDL", "0")
    plain = _layer(N, M, 0.04, 1)
    monkeypatch.setenv("LCT_PDL", "1")
    pdl = _layer(N, M, 0.04, 1)
    xs = [torch.rand(B, 1, M, N, N, device="cuda") for _ in range(4)]
    gs = [torch.randn(B, 1, M, N, N, device="cuda") for _ in range(4)]
    tb, te = [0] * B, [M] * B
    with torch.no_grad():
        want = [(plain(x, tb, te).clone(), plain._plan.backward(g, tb, te, M).clone()) for x, g in zip(xs, gs)]
        torch.cuda.synchronize()
        for _ in range(5):
            got = []
            for x, g in zip(xs, gs):          # no synchronisation between the calls; torch's caching allocator hands
                got.append((pdl._plan.forward(x, tb, te), pdl._plan.backward(g, tb, te, M)))   # every call the same workspace block
            torch.cuda.synchronize()
            for (y, gx), (wy, wgx) in zip(got, want):
                assert torch.equal(y, wy) and torch.equal(gx, wgx)


def test_numpy_path_of_the_reference(parity_log):
    """the CUDA volume against the outputs of /root/reference/utils/lct.py itself
    (tests/golden/make_golden_numpy_lct.py): the un-clamped volume (lct.py:41-59) and the three displayed views."""
    z, meas = numpy_lct_fixture()
    N, M = int(z["N"]), int(z["M"])
    layer = _layer(N, M, float(z["bin_len"]), 1)
    x = torch.from_numpy(np.ascontiguousarray(np.transpose(meas, [2, 0, 1]))).view(1, 1, M, N, N)
    vol = layer(x.cuda(), [0], [M]).cpu().numpy().reshape(M, N, N)
    e = O.rel_l2(vol, z["volume"])
    views = O.display_views(vol)
    ev = max(O.rel_l2(views[k], z[k]) for k in ("front", "left", "top"))
    parity_log(M=M, N=N, volume_rel_l2=e, views_rel_l2=ev, against="utils/lct.py golden")
    assert e <= TOL_Y and ev <= TOL_Y


def test_non_float32_input_is_refused():
    layer = _layer(8, 32, 0.16, 1)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 1, 32, 8, 8, device="cuda", dtype=torch.float64), [0], [32])


def test_stream_groups_do_not_change_results(monkeypatch):
    """The two-stream channel split (default) returns the same bits as the single-stream order."""
    import hiddenpose_b200 as hp
    M, N, B = 64, 32, 5
    x = torch.rand(B, 1, M, N, N, device="cuda")
    outs = []
    for groups in ("1", "2", "3"):
        monkeypatch.setenv("LCT_STREAM_GROUPS", groups)
        layer = hp.lct(spatial=N, crop=M, bin_len=0.08)
        layer.todev("cuda:0", 1)
        outs.append(layer(x, [0] * B, [M] * B))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


# the full compiled grid.  The oracle's constructor is the reference's host code (definePsf + np.fft.fftn of the
# padded volume: 11 s at 8 Mi voxels, 46 s at 32 Mi), so the two largest shapes -- (256, 256) and (512, 256) -- are
# compared with the reference's own outputs instead (test_golden_forward_backward: m256n256_full, m512n256_full).
_SIZES = [(M, N) for M in (32, 64, 128, 256, 512) for N in (8, 16, 32, 64, 128, 256) if M * N * N <= (1 << 23)]


def test_every_compiled_size_has_a_forward_and_gradient_comparison():
    """Bookkeeping: each (M, N) the library compiles is covered by the oracle grid below or by a golden fixture."""
    golden = {case_shape(n) for n in case_names()}
    for M in (32, 64, 128, 256, 512):
        for N in (8, 16, 32, 64, 128, 256):
            assert (M, N) in _SIZES or (M, N) in golden, (M, N)
    for shape in ((256, 64), (512, 128), (128, 128), (512, 256)):          # BASELINE.json configs
        assert shape in golden


@pytest.mark.parametrize("M,N", _SIZES)
def test_every_compiled_size_forward_backward(M, N, parity_log):
    """Every (time_bins, spatial) instantiation the library compiles, forward and backward, with a
    partial window (and two channels up to 2 Mi voxels, one above), against the oracle's op sequence (run
    with torch's CUDA FFT for speed: test-only; the volumes are too many for the CPU oracle in one suite)."""
    bl = 0.01 * 512 / M
    D = 2 if M * N * N <= (1 << 21) else 1
    layer = _layer(N, M, bl, D)
    orc = O.LctOracle(N, M, bl)
    gen = torch.Generator(device="cuda").manual_seed(M * 1000 + N)
    tin = M - 3
    x = torch.rand(1, D, tin, N, N, device="cuda", generator=gen)
    g = torch.randn(1, D, M, N, N, device="cuda", generator=gen)
    xd = x.clone().requires_grad_(True)
    y = layer(xd, [2], [2 + tin])
    y.backward(g)
    xo = x.clone().requires_grad_(True)
    yo = orc.forward(xo, [2], [2 + tin])
    yo.backward(g)
    ey, eg = O.rel_l2(y.detach().cpu(), yo.detach().cpu()), O.rel_l2(xd.grad.cpu(), xo.grad.cpu())
    parity_log(M=M, N=N, volume_rel_l2=ey, grad_rel_l2=eg, against="oracle op sequence on torch CUDA")
    assert ey <= TOL_Y
    assert eg <= TOL_G


def test_cuda_graph_capture_replays_the_layer():
    """The whole forward (two internal streams, no allocation or sync inside the library) can be captured
    in a CUDA graph and replayed on new data -- what a launch-bound caller with tiny batches would do."""
    M, N, B = 64, 16, 2
    layer = _layer(N, M, 0.08, 1)
    static_x = torch.rand(B, 1, M, N, N, device="cuda")
    with torch.no_grad():
        layer(static_x, [0] * B, [M] * B)                       # warm up outside the capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_y = layer(static_x, [0] * B, [M] * B)
        for _ in range(2):
            fresh = torch.rand_like(static_x)
            static_x.copy_(fresh)
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(static_y, layer(fresh, [0] * B, [M] * B))


def test_lct_graph_replays_layer_and_normalize():
    """hiddenpose_b200.LctGraph: one-launch replay (layer, or layer + normalize_feature) equals the eager calls."""
    import hiddenpose_b200 as hp
    M, N = 64, 32
    fp = hp.FeaturePropagation(image_size=N, time_size=M, bin_len=0.08, wall_size=2.0, dnum=1, dev="cuda:0")
    x = torch.rand(1, 1, M - 4, N, N, device="cuda")
    tb, te = [2], [M - 2]
    g_plain = hp.LctGraph(fp, x.shape, tb, te)
    g_norm = hp.LctGraph(fp, x.shape, tb, te, normalize=True)
    with torch.no_grad():
        for _ in range(2):
            x = torch.rand_like(x)
            want = fp(x, tb, te)
            assert torch.equal(g_plain(x), want)
            assert torch.equal(g_norm(x), hp.normalize_feature(fp(x, tb, te)))
    with pytest.raises(ValueError):
        g_plain(torch.rand(2, 1, M - 4, N, N, device="cuda"))


@pytest.mark.parametrize("M,N,C", [(32, 8, 3), (32, 16, 5), (64, 64, 2), (32, 128, 2), (32, 256, 1)])
def test_bp_laplacian_tail_and_its_transpose(M, N, C):
    """method == 'bp' (tflct.py:164-175): ReplicationPad3d(2) -> conv3d with the 5x5x5 filter -> zero time slice 0.
    The tiled stencil kernel against torch's own pad + conv3d, and its transpose (tiled interior + boundary shell)
    against autograd of that expression; every tile shape the layer can be built with."""
    layer = _layer(N, M, 0.04, 1, method="bp")
    plan = layer._plan
    w = torch.from_numpy(np.asarray(plan.lapw, np.float32).reshape(1, 1, 5, 5, 5)).cuda()
    vol = torch.randn(C, 1, M, N, N, device="cuda", requires_grad=True)
    ref = torch.nn.functional.conv3d(torch.nn.functional.pad(vol, (2,) * 6, mode="replicate"), w)
    ref = torch.cat([torch.zeros_like(ref[:, :, :1]), ref[:, :, 1:]], dim=2)
    got = plan.laplacian(vol.detach().view(C, 1, M, N, N), adjoint=False)
    assert O.rel_l2(got.cpu(), ref.detach().cpu()) <= 2e-6
    assert torch.count_nonzero(got[:, :, 0]) == 0
    g = torch.randn_like(ref)
    ref.backward(g)
    gv = plan.laplacian(g, adjoint=True)
    assert O.rel_l2(gv.cpu(), vol.grad.cpu()) <= 2e-6


@pytest.mark.parametrize("M,N,method", [(64, 16, "lct"), (128, 64, "lct"), (64, 128, "lct"), (64, 32, "bp")])
def test_device_built_filter_matches_host_built(M, N, method, monkeypatch):
    """Row f3: the filter built on the GPU from the PSF support gives the same volumes as the one built on
    the host with NumPy in double precision (both fused and five-kernel layouts)."""
    import hiddenpose_b200 as hp
    x = torch.rand(2, 1, M, N, N, device="cuda")
    outs = []
    for host in ("0", "1"):
        monkeypatch.setenv("HIDDENPOSE_LCT_HOST_FILTER", host)
        layer = hp.lct(spatial=N, crop=M, bin_len=0.01 * 512 / M, method=method)
        layer.todev("cuda:0", 1)
        outs.append(layer(x, [0, 0], [M, M]))
    assert O.rel_l2(outs[0].cpu(), outs[1].cpu()) <= 2e-6
