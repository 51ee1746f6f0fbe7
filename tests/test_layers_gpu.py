"""Single-layer parity on the GPU: the CUDA conv kernels, called through the C ABI test entry points,
against the known-answer vectors dumped from the reference layer classes
(layers/conv_layer.py CausalConv1d / CausalConvTranspose1d .inference; tests/golden/layers.npz)."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(g, name):
    return {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def lib():
    from audiodec_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize("name", ["conv_k7_d3", "conv_k6_s3", "conv_k10_s5_ragged", "conv_k11_d5_g3", "conv_short_chunk"])
def test_causal_conv_matches_reference(lib, golden_dir, name, conv_path):
    from audiodec_b200 import _lib
    c = _case(np.load(os.path.join(golden_dir, "layers.npz")), name)
    cin, cout, k, s, d, grp, T = [int(v) for v in c["cfg"]]
    state = np.zeros((1, cin, (k - 1) * d), np.float32)
    w, b = np.ascontiguousarray(c["w"]), np.ascontiguousarray(c["b"])
    for xi, yi in (("x0", "y0"), ("x1", "y1")):      # two consecutive chunks exercise the state carry
        x = np.ascontiguousarray(c[xi])
        y = np.zeros_like(c[yi])
        rc = lib.adec_test_causal_conv(0, _p(x), 1, cin, T, _p(w), _p(b), cout, k, s, d, grp, 0, 0.0, _p(state), _p(y))
        assert rc == 0, _lib.last_error(None)
        np.testing.assert_allclose(y, c[yi], atol=2e-5, rtol=0)
        xx = np.concatenate([np.zeros_like(state) if xi == "x0" else prev, x], -1)
        np.testing.assert_array_equal(state, xx[:, :, xx.shape[-1] - state.shape[-1]:])   # conv_layer.py:155
        prev = state.copy()


@pytest.mark.parametrize("name", ["convtr_s5", "convtr_s3"])
def test_causal_convtr_matches_reference(lib, golden_dir, name, conv_path):
    from audiodec_b200 import _lib
    c = _case(np.load(os.path.join(golden_dir, "layers.npz")), name)
    cin, cout, k, s, _, _, T = [int(v) for v in c["cfg"]]
    state = np.zeros((1, cin, 1), np.float32)
    w, b = np.ascontiguousarray(c["w"]), np.ascontiguousarray(c["b"])
    for xi, yi in (("x0", "y0"), ("x1", "y1")):
        x = np.ascontiguousarray(c[xi])
        y = np.zeros_like(c[yi])
        rc = lib.adec_test_causal_convtr(0, _p(x), 1, cin, T, _p(w), _p(b), cout, s, _p(state), _p(y))
        assert rc == 0, _lib.last_error(None)
        np.testing.assert_allclose(y, c[yi], atol=2e-5, rtol=0)
        np.testing.assert_array_equal(state, x[:, :, -1:])


def test_conv_with_preactivation_and_batch(lib):
    """ELU / LeakyReLU pre-activation + batch of 3 streams against torch fp32 on CPU."""
    from audiodec_b200 import _lib
    torch.manual_seed(0)
    for act, slope in ((1, 0.0), (2, 0.1)):
        cin, cout, k, d, T, B = 32, 64, 7, 9, 300, 3
        x = torch.randn(B, cin, T)
        w = torch.randn(cout, cin, k) / (cin * k) ** 0.5
        st = torch.randn(B, cin, (k - 1) * d)
        f = torch.nn.functional.elu if act == 1 else (lambda v: torch.nn.functional.leaky_relu(v, slope))
        ref = torch.nn.functional.conv1d(torch.cat([st, f(x)], -1), w, None, dilation=d)
        xn, wn, stn = x.numpy().copy(), w.numpy().copy(), st.numpy().copy()
        y = np.zeros((B, cout, T), np.float32)
        rc = lib.adec_test_causal_conv(0, _p(xn), B, cin, T, _p(wn), None, cout, k, 1, d, 1, act, slope, _p(stn), _p(y))
        assert rc == 0, _lib.last_error(None)
        np.testing.assert_allclose(y, ref.numpy(), atol=2e-5, rtol=0)
        np.testing.assert_allclose(stn, torch.cat([st, f(x)], -1)[:, :, -(k - 1) * d:].numpy(), atol=2e-6)


@pytest.mark.parametrize("C,d,T,B", [(32, 1, 700, 2), (64, 3, 300, 1), (64, 9, 130, 2), (128, 9, 260, 1)])
def test_residual_unit_fused(lib, C, d, T, B, conv_path):
    """Fused residual unit (residual_unit.py:78-81) against torch fp32 on CPU, two consecutive chunks."""
    from audiodec_b200 import _lib
    torch.manual_seed(C + d)
    w1 = torch.randn(C, C, 7) * 1.2 / (7 * C) ** 0.5
    w2 = torch.randn(C, C, 1) * 0.6 / C ** 0.5
    st = torch.zeros(B, C, 6 * d)
    stn = st.numpy().copy()
    for _ in range(2):
        x = torch.randn(B, C, T)
        xx = torch.cat([st, torch.nn.functional.elu(x)], -1)
        st = xx[:, :, -6 * d:]
        mid = torch.nn.functional.conv1d(xx, w1, None, dilation=d)
        ref = x + torch.nn.functional.conv1d(torch.nn.functional.elu(mid), w2)
        y = np.zeros((B, C, T), np.float32)
        rc = lib.adec_test_residual_unit(0, _p(x.numpy().copy()), B, C, T, _p(w1.numpy().copy()), _p(w2.numpy().copy()), 7, d, _p(stn), _p(y))
        assert rc == 0, _lib.last_error(None)
        np.testing.assert_allclose(y, ref.numpy(), atol=1e-5, rtol=0)
        np.testing.assert_allclose(stn, st.numpy(), atol=1e-6)


@pytest.mark.parametrize("cin,cout,k,s,d", [(128, 256, 7, 1, 1), (256, 128, 3, 1, 1), (64, 128, 8, 4, 1), (256, 512, 10, 5, 1), (512, 64, 3, 1, 1)])
def test_wide_convs_multi_piece(lib, cin, cout, k, s, d):
    """Convs whose input is wider than one 32-channel piece / one output tile, vs torch fp32 CPU (tight tolerance:
    this is the check that the tensor-core path is fp32-grade, not TF32-grade)."""
    from audiodec_b200 import _lib
    torch.manual_seed(cin + cout)
    T, B = 310, 2
    x = torch.randn(B, cin, T)
    w = torch.randn(cout, cin, k) / (cin * k) ** 0.5
    bias = torch.randn(cout) * 0.1
    st = torch.randn(B, cin, (k - 1) * d)
    ref = torch.nn.functional.conv1d(torch.cat([st, x], -1), w, bias, stride=s, dilation=d)
    stn = st.numpy().copy()
    y = np.zeros(tuple(ref.shape), np.float32)
    rc = lib.adec_test_causal_conv(0, _p(x.numpy().copy()), B, cin, T, _p(w.numpy().copy()), _p(bias.numpy().copy()), cout, k, s, d, 1, 0, 0.0,
                                   _p(stn), _p(y))
    assert rc == 0, _lib.last_error(None)
    np.testing.assert_allclose(y, ref.numpy(), atol=3e-6, rtol=0)
