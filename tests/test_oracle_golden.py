"""The oracle (oracle/audiodec_oracle.py) against the golden vectors dumped from the
unmodified reference (tests/golden/make_golden.py).  CPU only.

Tolerances: the oracle issues the same torch CPU ops as the reference, so in the build
container the match is bit-exact; on another host CPU (the GPU box) oneDNN/MKL may pick
different kernels, so waveforms are compared at 2e-5 max-abs and indices must still be
bit-identical (the goldens' smallest relative top-2 margin is 6.9e-5, far above conv
rounding noise)."""
import os

import numpy as np
import pytest
import torch

from audiodec_b200 import synthetic as S
from oracle import audiodec_oracle as O

TOL = 2e-5


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_weights_regenerate_bit_identically(golden_dir, symad_sd, hifigan_sd):
    g = _load(golden_dir, "v1_oneshot.npz")
    assert S.state_dict_digest(symad_sd) == str(g["enc_digest"])
    assert S.state_dict_digest(hifigan_sd) == str(g["dec_digest"])


def test_symad_oneshot(golden_dir, symad_sd):
    g = _load(golden_dir, "symad_oneshot.npz")
    fresh = O.SymADOracle(S.SYMAD_PARAMS, symad_sd)
    zq0 = fresh.initial_encoder(8192)
    assert zq0.shape == (1, 28, 64)
    np.testing.assert_allclose(zq0.numpy(), g["warm_zq"], atol=TOL)
    c = O.CodecOracle(S.SYMAD_PARAMS, symad_sd)
    z, idx, zq, y = c.run(torch.from_numpy(g["x"]))
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (8, 40)
    np.testing.assert_array_equal(idx.numpy(), g["idx"])
    np.testing.assert_allclose(z.numpy(), g["z"], atol=TOL)
    np.testing.assert_allclose(zq.numpy(), g["zq"], atol=TOL)
    np.testing.assert_allclose(y.numpy(), g["y"], atol=TOL)


def test_symad_stream_chunks(golden_dir, symad_sd):
    g = _load(golden_dir, "symad_stream.npz")
    c = O.CodecOracle(S.SYMAD_PARAMS, symad_sd)
    x = torch.from_numpy(g["x"])
    n = int(g["chunk"])
    outs = [c.run(x[:, :, i:i + n]) for i in range(0, x.shape[-1], n)]
    np.testing.assert_array_equal(torch.cat([o[1] for o in outs], -1).numpy(), g["idx"])
    np.testing.assert_allclose(torch.cat([o[3] for o in outs], -1).numpy(), g["y"], atol=TOL)
    # chunk invariance: streamed == one-shot
    one = _load(golden_dir, "symad_oneshot.npz")
    np.testing.assert_array_equal(g["idx"], one["idx"])
    np.testing.assert_allclose(g["y"], one["y"], atol=5e-6)


def test_symad_ragged_length(golden_dir, symad_sd):
    g = _load(golden_dir, "symad_ragged.npz")
    c = O.CodecOracle(S.SYMAD_PARAMS, symad_sd)
    z, idx, zq, y = c.run(torch.from_numpy(g["x"]))
    assert z.shape[-1] == 14 and y.shape[-1] == 4200      # floor((T-1)/s)+1 per stage, T=4001
    np.testing.assert_array_equal(idx.numpy(), g["idx"])
    np.testing.assert_allclose(y.numpy(), g["y"], atol=TOL)


def test_symad_batch3(golden_dir, symad_sd):
    g = _load(golden_dir, "symad_batch3.npz")
    c = O.CodecOracle(S.SYMAD_PARAMS, symad_sd)
    z, idx, zq, y = c.run(torch.from_numpy(g["x"]))
    assert tuple(idx.shape) == (8, 3, 20) and tuple(zq.shape) == (3, 20, 64)
    np.testing.assert_array_equal(idx.numpy(), g["idx"])
    np.testing.assert_allclose(zq.numpy(), g["zq"], atol=TOL)
    np.testing.assert_allclose(y.numpy(), g["y"], atol=TOL)


def test_v1_vocoder(golden_dir, symad_sd, hifigan_sd):
    g = _load(golden_dir, "v1_oneshot.npz")
    c = O.CodecOracle(S.SYMAD_PARAMS, symad_sd, S.HIFIGAN_V1_PARAMS, hifigan_sd)
    z, idx, zq, y = c.run(torch.from_numpy(g["x"]))
    np.testing.assert_array_equal(idx.numpy(), g["idx"])
    np.testing.assert_allclose(y.numpy(), g["y"], atol=TOL)
    gs = _load(golden_dir, "v1_stream.npz")
    c = O.CodecOracle(S.SYMAD_PARAMS, symad_sd, S.HIFIGAN_V1_PARAMS, hifigan_sd)
    x = torch.from_numpy(gs["x"])
    ys = [c.run(x[:, :, i:i + 1500])[3] for i in range(0, x.shape[-1], 1500)]
    np.testing.assert_allclose(torch.cat(ys, -1).numpy(), gs["y"], atol=TOL)


def _case(g, name):
    return {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}


@pytest.mark.parametrize("name", ["conv_k7_d3", "conv_k6_s3", "conv_k10_s5_ragged", "conv_k11_d5_g3", "conv_short_chunk"])
def test_layer_causal_conv(golden_dir, name):
    c = _case(_load(golden_dir, "layers.npz"), name)
    cin, cout, k, s, d, grp, T = [int(v) for v in c["cfg"]]
    state = torch.zeros(1, cin, (k - 1) * d)
    w, b = torch.from_numpy(c["w"]), torch.from_numpy(c["b"])
    for xi, yi in (("x0", "y0"), ("x1", "y1")):
        y, state = O.causal_conv1d_infer(torch.from_numpy(c[xi]), w, b, state, s, d, grp)
        assert y.shape[-1] == (T - 1) // s + 1
        np.testing.assert_allclose(y.numpy(), c[yi], atol=1e-6)


@pytest.mark.parametrize("name", ["convtr_s5", "convtr_s3"])
def test_layer_causal_convtr(golden_dir, name):
    c = _case(_load(golden_dir, "layers.npz"), name)
    cin, cout, k, s, _, _, T = [int(v) for v in c["cfg"]]
    state = torch.zeros(1, cin, 1)
    w, b = torch.from_numpy(c["w"]), torch.from_numpy(c["b"])
    for xi, yi in (("x0", "y0"), ("x1", "y1")):
        x = torch.from_numpy(c[xi])
        prev = state
        y, state = O.causal_convtr1d_infer(x, w, b, state, s)
        np.testing.assert_allclose(y.numpy(), c[yi], atol=1e-6)
        # two-tap closed form: y[j*s+r] = b + W[:,:,r]^T x[j] + W[:,:,s+r]^T x[j-1]
        xx = torch.cat((prev, x), -1)[0]
        for j in range(T):
            for r in range(s):
                ref = b + w[:, :, r].T @ xx[:, j + 1] + w[:, :, s + r].T @ xx[:, j]
                np.testing.assert_allclose(ref.numpy(), c[yi][0, :, j * s + r], atol=2e-6)


def test_layer_rvq(golden_dir):
    c = _case(_load(golden_dir, "layers.npz"), "rvq")
    embeds = [torch.from_numpy(e) for e in c["embeds"]]
    zq, idx = O.rvq_forward_index(torch.from_numpy(c["x"]), embeds, True)
    np.testing.assert_array_equal(idx.squeeze(1).numpy(), c["idx"])
    np.testing.assert_allclose(zq.numpy(), c["zq"], atol=1e-6)
    cb = torch.stack([e.T for e in embeds]).reshape(-1, 16)
    np.testing.assert_allclose(O.rvq_lookup(idx.squeeze(1), cb).numpy(), c["lookup"], atol=1e-6)


VARIANTS = {   # golden file -> (encoder params, vocoder params or None)
    "v2_oneshot.npz": ("SYMAD_PARAMS", "HIFIGAN_V2_PARAMS"),
    "v0_oneshot.npz": ("SYMAD_PARAMS", "HIFIGAN_V0_PARAMS"),
    "aad_oneshot.npz": ("SYMAAD_PARAMS", None),
    "c16_oneshot.npz": ("SYMAD_C16_PARAMS", None),
}


@pytest.mark.parametrize("fname", sorted(VARIANTS))
def test_released_variants(golden_dir, fname):
    """HiFi-GAN v2 (k=3) / v0 (MultiReceptiveField), symAAD (weight-normed, ELU/tanh), 16-codebook hop-320 symAD:
    one-shot and two-chunk streaming vs the unmodified reference."""
    g = _load(golden_dir, fname)
    ep, vp = (getattr(S, n) if n else None for n in VARIANTS[fname])
    esd = S.symad_state_dict(ep, seed=0)
    vsd = S.hifigan_state_dict(vp, seed=1) if vp else None
    x = torch.from_numpy(g["x"])
    c = O.CodecOracle(ep, esd, vp, vsd)
    z, idx, zq, y = c.run(x)
    assert tuple(idx.shape) == tuple(g["idx"].shape)
    np.testing.assert_array_equal(idx.numpy(), g["idx"])
    np.testing.assert_allclose(y.numpy(), g["y"], atol=TOL)
    c = O.CodecOracle(ep, esd, vp, vsd)
    outs = [c.run(x[:, :, i:i + 3200]) for i in (0, 3200)]
    np.testing.assert_array_equal(torch.cat([o[1] for o in outs], -1).numpy(), g["idx_chunks"])
    np.testing.assert_allclose(torch.cat([o[3] for o in outs], -1).numpy(), g["y_chunks"], atol=TOL)


def test_bitstream_oracle_known_answer_and_round_trip():
    """oracle/bitstream_oracle.py against an independent statement of the same format (one little-endian big integer per
    frame) and its own inverse; frame sizes of the released configs."""
    from oracle import bitstream_oracle as BO
    assert BO.index_bits(1024) == 10 and BO.index_bits(1000) == 10 and BO.index_bits(2) == 1
    assert BO.frame_bytes(8, 1024) == 10 and BO.frame_bytes(16, 1024) == 20 and BO.frame_bytes(3, 1024) == 4
    rng = np.random.default_rng(0)
    for nq, n in ((8, 1024), (16, 1024), (3, 1000), (5, 7)):
        idx = rng.integers(0, n, (nq, 2, 6)) + n * np.arange(nq)[:, None, None]
        packed = BO.pack_indices(idx, n)
        bits = BO.index_bits(n)
        for b in range(2):
            for f in range(6):
                big = 0
                for i in range(nq):
                    big |= int(idx[i, b, f] - i * n) << (bits * i)
                assert bytes(packed[b, f]) == big.to_bytes(BO.frame_bytes(nq, n), "little")
        np.testing.assert_array_equal(BO.unpack_indices(packed, nq, n), idx)
    # hand-computed vector: local indices 1..8, 10 bits each
    idx = (np.arange(8) + 1 + 1024 * np.arange(8)).reshape(8, 1, 1)
    assert BO.pack_indices(idx, 1024).ravel().tolist() == [1, 8, 48, 0, 1, 5, 24, 112, 0, 2]


OFFLINE = {   # golden file -> (encoder params, vocoder params or None)
    "offline_symad.npz": ("SYMAD_PARAMS", None), "offline_aad.npz": ("SYMAAD_PARAMS", None),
    "offline_c16.npz": ("SYMAD_C16_PARAMS", None), "offline_v1.npz": ("SYMAD_PARAMS", "HIFIGAN_V1_PARAMS"),
    "offline_v0.npz": ("SYMAD_PARAMS", "HIFIGAN_V0_PARAMS"),
}


@pytest.mark.parametrize("fname", sorted(OFFLINE))
def test_offline_forward_oracle_vs_reference(golden_dir, fname):
    """the non-streaming batch forward of codecTest.py:78-95 (zero left-pad, replication pad on the
    transposed convs), oracle restatement vs vectors dumped from the reference's base Generator classes."""
    g = np.load(os.path.join(golden_dir, fname))
    ep, vp = (getattr(S, n) if n else None for n in OFFLINE[fname])
    enc = O.SymADOracle(ep, S.symad_state_dict(ep, seed=0))
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        z = enc.forward_encode(x)
        zq = enc.forward_quantize(z)
        if vp is None:
            y = O.SymADOracle(ep, S.symad_state_dict(ep, seed=0)).forward_decode(zq)
        else:
            y = O.HiFiGANOracle(vp, S.hifigan_state_dict(vp, seed=1)).forward(zq)
    # fp32 conv summation order depends on the host thread count (goldens: 4 threads), hence not bit-exact
    np.testing.assert_allclose(z.numpy(), g["z"], atol=5e-6)
    np.testing.assert_allclose(zq.numpy(), g["zq"], atol=5e-6)
    np.testing.assert_allclose(y.numpy(), g["y"], atol=1e-5)
    # and the streaming state of the oracle object is untouched (zeros) afterwards
    assert all(float(v.abs().max()) == 0.0 for v in enc.state.values())


def test_bitstream_and_shard_properties_hypothesis():
    """Size-independent properties: unpack(pack(idx)) == idx for any codebook count / size, frames are independent
    (packing a batch == packing its frames one by one), and shard bounds always tile [0, n) contiguously and evenly."""
    from hypothesis import given, settings, strategies as st
    from audiodec_b200.shard import shard_bounds
    from oracle import bitstream_oracle as BO

    @settings(max_examples=40, deadline=None)
    @given(nq=st.integers(1, 16), n=st.integers(2, 4096), b=st.integers(1, 3), f=st.integers(1, 5), seed=st.integers(0, 2**31 - 1))
    def bitstream(nq, n, b, f, seed):
        rng = np.random.default_rng(seed)
        idx = rng.integers(0, n, (nq, b, f)) + n * np.arange(nq)[:, None, None]
        p = BO.pack_indices(idx, n)
        assert p.shape == (b, f, BO.frame_bytes(nq, n)) and p.dtype == np.uint8
        np.testing.assert_array_equal(BO.unpack_indices(p, nq, n), idx)
        np.testing.assert_array_equal(p[b - 1, f - 1], BO.pack_indices(idx[:, b - 1:, f - 1:], n)[0, 0])
        pad_bits = 8 * BO.frame_bytes(nq, n) - nq * BO.index_bits(n)
        assert 0 <= pad_bits < 8 and int(p[0, 0, -1]) >> (8 - pad_bits) == 0 if pad_bits else True

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(0, 5000), w=st.integers(1, 16))
    def shards(n, w):
        bnd = shard_bounds(n, w)
        sizes = [e - s for s, e in bnd]
        assert len(bnd) == w and bnd[0][0] == 0 and bnd[-1][1] == n and sum(sizes) == n
        assert all(bnd[i][1] == bnd[i + 1][0] for i in range(w - 1)) and max(sizes) - min(sizes) <= 1
        assert sizes == sorted(sizes, reverse=True)

    bitstream()
    shards()
