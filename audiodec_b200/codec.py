"""Host-side mirrors of the reference's streaming generators, backed by the C-ABI library.

``SymADStreamGenerator`` stands in for ``models/autoencoder/AudioDec.py:166 StreamGenerator`` and
``HiFiGANStreamGenerator`` for ``models/vocoder/HiFiGAN.py:222 StreamGenerator``: same constructor
keywords (``config.yml`` ``generator_params``), same methods with the same tensor shapes
(``load_state_dict / eval / to / initial_encoder / initial_decoder / encode / quantize / lookup /
decode / reset_buffer``), so the objects can be returned from ``AudioCodec._load_encoder /
_load_decoder`` (bin/stream.py:38-45) unchanged.  All arithmetic happens in hand-written sm_100a
kernels (audiodec_b200/csrc); torch only provides device memory and the current stream.

Differences from the reference, all extensions:
  * batches: the reference's streaming state is (1,C,P) so only B=1 works (layers/conv_layer.py:144-146);
    here a batch of B independent streams is allowed.  A handle warmed with one stream replicates its
    state when first called with B>1.  ``quantize`` then returns (Nq,B,F) (what
    ``ResidualVQ.forward_index`` yields before its ``squeeze(1)``, vq_module.py:148-149) and ``lookup``
    returns (B,F,D).
  * there is no CPU path: ``.to('cpu')`` raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _check(rc, handle):
    if rc != 0:
        raise RuntimeError("audiodec_b200: " + _lib.last_error(handle))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _fill(arr, values):
    for i, v in enumerate(values):
        arr[i] = int(v)


class _StreamGeneratorBase:
    """Common plumbing: deferred handle creation (weights arrive before the device is known, exactly
    like ``Generator(**params)`` -> ``load_state_dict`` -> ``.to(device)`` in the reference)."""

    def __init__(self):
        self._lib = _lib.load()
        self._cfg = _lib.AdecConfig()
        self._sd = None
        self._h = None
        self._device = None

    # -- torch.nn.Module look-alikes ------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        self._sd = {k: v.detach().to(torch.float32).cpu().contiguous() for k, v in state_dict.items()}
        return self

    def eval(self):
        return self

    def to(self, device):
        if isinstance(device, torch.dtype):
            return self._set_dtype(device)
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("audiodec_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self._sd is None:
            raise RuntimeError("load_state_dict must be called before .to(device)")
        if self._h is not None:
            if device.index not in (None, self._device.index):
                raise RuntimeError("handle already lives on " + str(self._device))
            return self
        index = device.index if device.index is not None else torch.cuda.current_device()
        self._device = torch.device("cuda", index)
        h = ctypes.c_void_p()
        rc = self._lib.adec_create(ctypes.byref(self._cfg), index, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("audiodec_b200: " + _lib.last_error(None))
        self._h = h
        for key, t in self._sd.items():
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            _check(self._lib.adec_set_tensor(h, key.encode(), _ptr(t), shape, t.dim()), h)
        _check(self._lib.adec_finalize(h), h)
        self._sd = None
        return self

    def _set_dtype(self, dtype):
        """`module.to(torch.bfloat16)` of the reference: only the HiFi-GAN vocoder has a reduced-precision mode (bf16 conv operands,
        fp32 accumulation and fp32 activations in HBM); the encoder / projector / RVQ must stay fp32-grade for bit-identical indices."""
        if dtype == torch.float32:
            want = 0
        elif dtype == torch.bfloat16 and self._cfg.model_type == _lib.MODEL_HIFIGAN:
            want = 1
        else:
            raise NotImplementedError(f"{type(self).__name__} has no {dtype} mode (fp32 everywhere; bf16 for the HiFi-GAN vocoder only)")
        if self._h is not None and want != self._cfg.compute_dtype:
            raise RuntimeError("set the compute dtype before .to(device): the weights are packed when the handle is created")
        self._cfg.compute_dtype = want
        return self

    def bfloat16(self):
        return self._set_dtype(torch.bfloat16)

    def float(self):
        return self._set_dtype(torch.float32)

    def __del__(self):
        try:
            if self._h is not None:
                self._lib.adec_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------------------------
    def _ready(self):
        if self._h is None:
            raise RuntimeError("call .to('cuda:N') before using the codec")

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _in(self, t, dtype=torch.float32):
        if t.device != self._device:
            raise RuntimeError(f"input is on {t.device}, codec is on {self._device}")
        return t.to(dtype).contiguous()

    @staticmethod
    def _want(t, what, dim, size):
        """torch raises RuntimeError on a wrong feature dimension; the C ABI takes no D argument, so check here."""
        if t.dim() != 3 or t.size(dim) != size:
            raise RuntimeError(f"audiodec_b200: {what}: expected a 3-D tensor with size {size} in dim {dim}, got {tuple(t.shape)}")

    def _batch(self, b):
        n = self._lib.adec_n_streams(self._h)
        if n != b:
            _check(self._lib.adec_set_streams(self._h, b), self._h)

    @property
    def n_streams(self):
        self._ready()
        return self._lib.adec_n_streams(self._h)

    @property
    def launch_count(self):
        return int(self._lib.adec_launch_count(self._h)) if self._h is not None else 0

    def profile(self, enable=True):
        """start/stop per-launch CUDA-event timing (adec_profile)."""
        self._ready()
        _check(self._lib.adec_profile(self._h, int(enable)), self._h)

    def profile_report(self):
        """[(op name, ms, algorithmic bytes)] for every launch recorded since profile(True)."""
        self._ready()
        buf = ctypes.create_string_buffer(1 << 20)
        _check(self._lib.adec_profile_report(self._h, buf, len(buf)), self._h)
        rows = []
        for line in buf.value.decode().splitlines():
            name, ms, nbytes = line.split("\t")
            rows.append((name, float(ms), float(nbytes)))
        return rows

    def range_error(self):
        """True if an activation left the fp16-split range of the default conv engine since the last call (synchronises)."""
        self._ready()
        rc = self._lib.adec_range_error(self._h, self._stream())
        if rc < 0:
            raise RuntimeError("audiodec_b200: " + _lib.last_error(self._h))
        return bool(rc)

    def reset_buffer(self):
        """AudioDec.py:250-256 / HiFiGAN.py:298-305."""
        self._ready()
        _check(self._lib.adec_reset(self._h, self._stream()), self._h)


class SymADStreamGenerator(_StreamGeneratorBase):
    """models/autoencoder/AudioDec.py:166-256."""

    def __init__(self, input_channels=1, output_channels=1, encode_channels=32, decode_channels=32, code_dim=64,
                 codebook_num=8, codebook_size=1024, bias=True, enc_ratios=(2, 4, 8, 16), dec_ratios=(16, 8, 4, 2),
                 enc_strides=(3, 4, 5, 5), dec_strides=(5, 5, 4, 3), mode="causal", codec="audiodec",
                 projector="conv1d", quantier="residual_vq", nonlinear_activation="ELU",
                 nonlinear_activation_params={}, use_weight_norm=False):
        super().__init__()
        assert mode == "causal", f"Mode {mode} does not support streaming!"       # models/utils.py:13-15
        if codec not in ("audiodec", "activate_audiodec"):
            raise NotImplementedError(f"Codec ({codec}) is not supported!")          # AudioDec.py:53-60
        if projector != "conv1d" or quantier != "residual_vq":
            raise NotImplementedError("only projector='conv1d', quantier='residual_vq' are built")
        if nonlinear_activation != "ELU" or nonlinear_activation_params:
            raise NotImplementedError("only ELU(alpha=1) residual units are built")
        c = self._cfg
        c.model_type = _lib.MODEL_SYMAD
        c.input_channels, c.output_channels = input_channels, output_channels
        c.encode_channels, c.decode_channels = encode_channels, decode_channels
        c.code_dim, c.codebook_num, c.codebook_size = code_dim, codebook_num, codebook_size
        c.bias = int(bias)
        c.n_enc, c.n_dec = len(enc_strides), len(dec_strides)
        _fill(c.enc_ratios, enc_ratios), _fill(c.enc_strides, enc_strides)
        _fill(c.dec_ratios, dec_ratios), _fill(c.dec_strides, dec_strides)
        c.use_weight_norm = int(use_weight_norm)
        c.codec_activate = int(codec == "activate_audiodec")
        self.input_channels = input_channels
        self.code_dim, self.codebook_num = code_dim, codebook_num

    # -- streaming API ---------------------------------------------------------------------------------
    def initial_encoder(self, receptive_length, device):
        """AudioDec.py:216-221: push `receptive_length` zeros through encode/quantize/lookup."""
        self._ready()
        z = self.encode(torch.zeros(1, self.input_channels, receptive_length, device=self._device))
        return self.lookup(self.quantize(z))

    def initial_decoder(self, zq):
        self.decode(zq)                                                              # AudioDec.py:224-225

    def encode(self, x):
        """(B,1,T) float -> z (B,code_dim,F)   (AudioDec.py:228-234)"""
        self._ready()
        if x.dim() != 3:
            raise RuntimeError("encode expects (batch, channel, length)")
        if x.size(1) != self.input_channels:
            x = x.reshape(-1, self.input_channels, x.size(-1))
        x = self._in(x)
        b, _, t = x.shape
        self._batch(b)
        f = self._lib.adec_frames_for(self._h, t)
        z = torch.empty(b, self.code_dim, f, device=self._device, dtype=torch.float32)
        _check(self._lib.adec_encode(self._h, _ptr(x), b, t, _ptr(z), self._stream()), self._h)
        return z

    def quantize(self, z):
        """z (B,code_dim,F) -> idx int64 (Nq,F) for B==1 else (Nq,B,F)   (AudioDec.py:237-239)"""
        self._ready()
        self._want(z, "quantize", 1, self.code_dim)
        z = self._in(z)
        b, _, f = z.shape
        idx = torch.empty(self.codebook_num, b, f, device=self._device, dtype=torch.int64)
        _check(self._lib.adec_quantize(self._h, _ptr(z), b, f, _ptr(idx), self._stream()), self._h)
        return idx.squeeze(1) if b == 1 else idx

    def quantize_fused(self, z, want_idx=True, want_packed=False, want_zq=True):
        """quantize -> [pack] -> lookup in ONE launch (adec_quantize_ex): returns (idx or None, packed or None, zq or None) with the
        shapes of quantize() / pack() / lookup().  Bit-identical to the three separate calls."""
        self._ready()
        self._want(z, "quantize_fused", 1, self.code_dim)
        z = self._in(z)
        b, _, f = z.shape
        idx = torch.empty(self.codebook_num, b, f, device=self._device, dtype=torch.int64) if want_idx else None
        packed = torch.empty(b, f, self.packed_frame_bytes(), device=self._device, dtype=torch.uint8) if want_packed else None
        zq = torch.empty(b, f, self.code_dim, device=self._device, dtype=torch.float32) if want_zq else None
        _check(self._lib.adec_quantize_ex(self._h, _ptr(z), b, f, _ptr(idx) if want_idx else None, _ptr(packed) if want_packed else None,
                                          _ptr(zq) if want_zq else None, self._stream()), self._h)
        if b == 1:
            idx = idx.squeeze(1) if idx is not None else None
            packed = packed.squeeze(0) if packed is not None else None
        return idx, packed, zq

    def lookup_packed(self, packed):
        """uint8 (F,bytes) -> zq (1,F,D); (B,F,bytes) -> (B,F,D): lookup straight from the bitstream (unpack fused into lookup)."""
        self._ready()
        packed = self._in(packed, torch.uint8)
        if packed.dim() == 2:
            packed = packed.unsqueeze(0)
        b, f, nb = packed.shape
        if nb != self.packed_frame_bytes():
            raise RuntimeError(f"audiodec_b200: lookup_packed: expected {self.packed_frame_bytes()} bytes per frame, got {nb}")
        zq = torch.empty(b, f, self.code_dim, device=self._device, dtype=torch.float32)
        _check(self._lib.adec_lookup_packed(self._h, _ptr(packed), b, f, _ptr(zq), self._stream()), self._h)
        return zq

    def lookup(self, idx):
        """idx (Nq,F) -> zq (1,F,D); (Nq,B,F) -> (B,F,D)   (AudioDec.py:242-243)"""
        self._ready()
        idx = self._in(idx, torch.int64)
        if idx.dim() == 2:
            idx = idx.unsqueeze(1)
        if idx.dim() != 3 or idx.size(0) != self.codebook_num:
            raise RuntimeError(f"audiodec_b200: lookup: expected ({self.codebook_num},F) or ({self.codebook_num},B,F) indices, got {tuple(idx.shape)}")
        _, b, f = idx.shape
        zq = torch.empty(b, f, self.code_dim, device=self._device, dtype=torch.float32)
        _check(self._lib.adec_lookup(self._h, _ptr(idx), b, f, _ptr(zq), self._stream()), self._h)
        return zq

    # ---- non-streaming batch forward (codecTest.py:78-95).  These calls discard the streaming state.
    def encode_offline(self, x):
        """x (B,1,T) -> z (B,code_dim,F): Encoder.forward + Projector.forward, zero left-pad (conv_layer.py:148-151)."""
        self._ready()
        if x.dim() != 3:
            raise RuntimeError("encode_offline expects (batch, channel, length)")
        if x.size(1) != self.input_channels:                 # same fold as encode(): every audio channel is its own batch row
            x = x.reshape(-1, self.input_channels, x.size(-1))
        x = self._in(x)
        b, _, t = x.shape
        f = self._lib.adec_frames_for(self._h, t)
        z = torch.empty(b, self.code_dim, f, device=self._device, dtype=torch.float32)
        _check(self._lib.adec_encode_offline(self._h, _ptr(x), b, t, _ptr(z), self._stream()), self._h)
        return z

    def quantize_offline(self, z):
        """z (B,code_dim,F) -> (zq (B,code_dim,F) channels-first like Quantizer.forward (quantizer.py:31-34), idx (Nq,B,F))."""
        self._ready()
        self._want(z, "quantize_offline", 1, self.code_dim)
        z = self._in(z)
        b, _, f = z.shape
        idx = torch.empty(self.codebook_num, b, f, device=self._device, dtype=torch.int64)
        _check(self._lib.adec_quantize(self._h, _ptr(z), b, f, _ptr(idx), self._stream()), self._h)
        zq = torch.empty(b, f, self.code_dim, device=self._device, dtype=torch.float32)
        _check(self._lib.adec_lookup(self._h, _ptr(idx), b, f, _ptr(zq), self._stream()), self._h)
        return zq.transpose(1, 2), idx

    def decode_offline(self, zq):
        """zq (B,code_dim,F) channels-first (what Decoder.forward takes, decoder.py:135-140) -> y (B,1,F*hop); transposed convs
        replicate their first input frame (conv_layer.py:189-192)."""
        return _decode_offline(self, zq)

    # ---- index bitstream (the reference queues the raw int64 tensor, bin/stream.py:224)
    def packed_frame_bytes(self):
        self._ready()
        return self._lib.adec_packed_frame_bytes(self._h)

    def pack(self, idx):
        """idx (Nq,F) -> uint8 (F,bytes); (Nq,B,F) -> (B,F,bytes): Nq x ceil(log2 N)-bit local indices per frame."""
        self._ready()
        idx = self._in(idx, torch.int64)
        two_d = idx.dim() == 2
        if two_d:
            idx = idx.unsqueeze(1)
        nq, b, f = idx.shape
        if nq != self.codebook_num:
            raise RuntimeError(f"audiodec_b200: pack: expected {self.codebook_num} index rows, got {nq}")
        out = torch.empty(b, f, self.packed_frame_bytes(), device=self._device, dtype=torch.uint8)
        _check(self._lib.adec_pack_indices(self._h, _ptr(idx), b, f, _ptr(out), self._stream()), self._h)
        return out.squeeze(0) if two_d else out

    def unpack(self, packed):
        """uint8 (F,bytes) -> idx (Nq,F); (B,F,bytes) -> (Nq,B,F) int64 flat indices, ready for lookup()."""
        self._ready()
        packed = self._in(packed, torch.uint8)
        two_d = packed.dim() == 2
        if two_d:
            packed = packed.unsqueeze(0)
        b, f, nb = packed.shape
        if nb != self.packed_frame_bytes():
            raise RuntimeError(f"audiodec_b200: unpack: expected {self.packed_frame_bytes()} bytes per frame, got {nb}")
        idx = torch.empty(self.codebook_num, b, f, device=self._device, dtype=torch.int64)
        _check(self._lib.adec_unpack_indices(self._h, _ptr(packed), b, f, _ptr(idx), self._stream()), self._h)
        return idx.squeeze(1) if two_d else idx

    def index_error(self):
        """True if lookup / pack / unpack met an out-of-range index since the last call (synchronises the stream)."""
        self._ready()
        rc = self._lib.adec_index_error(self._h, self._stream())
        if rc < 0:
            raise RuntimeError("audiodec_b200: " + _lib.last_error(self._h))
        return bool(rc)

    def decode(self, zq):
        """zq (B,F,D) channels-last -> y (B,1,F*hop)   (AudioDec.py:246-247)"""
        self._ready()
        self._want(zq, "decode", 2, self.code_dim)
        zq = self._in(zq)
        b, f, _ = zq.shape
        self._batch(b)
        y = torch.empty(b, 1, f * self._lib.adec_hop_length(self._h), device=self._device, dtype=torch.float32)
        _check(self._lib.adec_decode(self._h, _ptr(zq), b, f, _ptr(y), self._stream()), self._h)
        return y


def _decode_offline(gen, zq):
    gen._ready()
    gen._want(zq, "decode_offline / forward", 1, getattr(gen, "code_dim", None) or gen.in_channels)
    zq = gen._in(zq)
    b, _, f = zq.shape
    zq_cl = zq.transpose(1, 2).contiguous()                       # the kernels' native channels-last (B,F,D)
    y = torch.empty(b, 1, f * gen._lib.adec_hop_length(gen._h), device=gen._device, dtype=torch.float32)
    _check(gen._lib.adec_decode_offline(gen._h, _ptr(zq_cl), b, f, _ptr(y), gen._stream()), gen._h)
    return y


class HiFiGANStreamGenerator(_StreamGeneratorBase):
    """models/vocoder/HiFiGAN.py:222-305 (AD v1: groups>1 and a single resblock kernel -> MultiGroupConv1d)."""

    def __init__(self, in_channels=80, out_channels=1, channels=512, kernel_size=7, upsample_scales=(8, 8, 2, 2),
                 upsample_kernel_sizes=(16, 16, 4, 4), resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilations=[(1, 3, 5), (1, 3, 5), (1, 3, 5)], groups=1, bias=True, use_additional_convs=True,
                 nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1},
                 use_weight_norm=True, stats=None):
        super().__init__()
        assert kernel_size % 2 == 1, "Kernel size must be odd number."               # HiFiGAN.py:73-75
        assert len(upsample_scales) == len(upsample_kernel_sizes)
        assert len(resblock_dilations) == len(resblock_kernel_sizes)
        multi_group = len(resblock_kernel_sizes) == 1 and groups > 1              # HiFiGAN.py:78-81
        if not multi_group:
            # MultiReceptiveField (AD v0): one residual block per kernel size, outputs averaged (multi_fusion.py:23-79)
            if groups != 1 or any(list(d) != list(resblock_dilations[0]) for d in resblock_dilations):
                raise NotImplementedError("MultiReceptiveField is built for groups=1 and identical dilation lists")
        if nonlinear_activation != "LeakyReLU" or not use_additional_convs or not bias:
            raise NotImplementedError("only LeakyReLU + additional convs + bias is built")
        c = self._cfg
        c.model_type = _lib.MODEL_HIFIGAN
        c.in_channels, c.out_channels, c.channels, c.kernel_size = in_channels, out_channels, channels, kernel_size
        c.n_up = len(upsample_scales)
        _fill(c.upsample_scales, upsample_scales), _fill(c.upsample_kernel_sizes, upsample_kernel_sizes)
        c.resblock_kernel_size = resblock_kernel_sizes[0]
        c.n_resblocks = 0 if multi_group else len(resblock_kernel_sizes)
        _fill(c.resblock_kernel_sizes, resblock_kernel_sizes)
        c.n_dil = len(resblock_dilations[0])
        _fill(c.resblock_dilations, resblock_dilations[0])
        c.groups = groups
        c.negative_slope = float(nonlinear_activation_params.get("negative_slope", 0.01))
        c.use_weight_norm = int(use_weight_norm)
        c.has_stats = int(stats is not None)       # mean/scale themselves come from the state dict (HiFiGAN.py:206-219)
        self.in_channels = in_channels

    def initial_decoder(self, c):
        self.decode(c)                                                                # HiFiGAN.py:264-265

    def decode(self, c):
        """zq (B,F,in_channels) channels-last -> y (B,1,F*prod(scales)) in (-1,1)   (HiFiGAN.py:268-296)"""
        self._ready()
        self._want(c, "decode", 2, self.in_channels)
        c = self._in(c)
        b, f, _ = c.shape
        self._batch(b)
        y = torch.empty(b, 1, f * self._lib.adec_hop_length(self._h), device=self._device, dtype=torch.float32)
        _check(self._lib.adec_decode(self._h, _ptr(c), b, f, _ptr(y), self._stream()), self._h)
        return y

    def forward(self, c):
        """Generator.forward (HiFiGAN.py:140-160), the non-streaming path: c (B,in_channels,F) channels-first -> y (B,1,F*hop).
        Discards the streaming state."""
        return _decode_offline(self, c)

    __call__ = forward


class OfflineCodec:
    """Mirror of codecTest.py's TestMain.encode / decode (codecTest.py:78-95): the non-streaming batch path.
    `encoder` is a SymADStreamGenerator, `decoder` a SymADStreamGenerator or HiFiGANStreamGenerator, both already on a
    CUDA device; they must not be the handles a live stream is using (offline calls reset the causal state)."""

    def __init__(self, encoder, decoder, multi_channel=False):
        if multi_channel and encoder.input_channels == 1:
            # codecTest.py:84-86 feeds (1,C,T) to a multi-channel generator; a mono generator would silently encode channel 0 only
            raise NotImplementedError("multi_channel=True needs a generator with input_channels > 1 (only mono generators are built)")
        self.encoder, self.decoder, self.multi_channel = encoder, decoder, multi_channel

    def encode(self, audio):
        """audio (T,C) float array -> zq (C,code_dim,F) (or (1,code_dim,F) when multi_channel)."""
        x = torch.as_tensor(audio, dtype=torch.float32).to(self.encoder._device)
        x = x.transpose(1, 0).unsqueeze(0) if self.multi_channel else x.transpose(1, 0).unsqueeze(1)
        zq, _ = self.encoder.quantize_offline(self.encoder.encode_offline(x))
        return zq

    def decode(self, zq):
        return self.decoder.forward(zq) if isinstance(self.decoder, HiFiGANStreamGenerator) else self.decoder.decode_offline(zq)


def codec_host(encoder: SymADStreamGenerator, decoder, x_host: torch.Tensor, want_idx=True, reuse_buffers=False):
    """Whole path on HOST buffers through ``adec_codec_host`` (H2D + encode + quantize + lookup + decode +
    D2H), i.e. what demoFile.py:55-62 does around the four calls.  x_host: (B,1,T) float32 CPU tensor
    (pinned for full PCIe speed).  With reuse_buffers the returned tensors are views of cached pinned buffers that
    the next call overwrites.  Returns (idx (Nq,B,F) int64 CPU or None, y (B,1,F*hop) float32 CPU)."""
    encoder._ready(), decoder._ready()
    assert x_host.device.type == "cpu" and x_host.dtype == torch.float32 and x_host.is_contiguous()
    b, _, t = x_host.shape
    encoder._batch(b)
    if decoder is not encoder:
        decoder._batch(b)
    lib = encoder._lib
    f = lib.adec_frames_for(encoder._h, t)
    hop = lib.adec_hop_length(decoder._h)
    pin = x_host.is_pinned()
    # page-locked result buffers are expensive to create (cudaHostAlloc): keep one set per shape and hand out views;
    # pass reuse_buffers=False to get fresh tensors that the next call will not overwrite
    cache = encoder.__dict__.setdefault("_host_out", {})
    key = (b, f, hop, pin, want_idx)
    if reuse_buffers and key in cache:
        idx, y = cache[key]
    else:
        idx = torch.empty(encoder.codebook_num, b, f, dtype=torch.int64, pin_memory=pin) if want_idx else None
        y = torch.empty(b, 1, f * hop, dtype=torch.float32, pin_memory=pin)
        if reuse_buffers:
            cache.clear()
            cache[key] = (idx, y)
    _check(lib.adec_codec_host(encoder._h, decoder._h, _ptr(x_host), b, t, _ptr(idx) if want_idx else None,
                               _ptr(y), encoder._stream()), encoder._h)
    return idx, y
