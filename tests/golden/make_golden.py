#!/usr/bin/env python3
"""Generate the golden vectors in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, CPU):

    python tests/golden/make_golden.py

It imports ``utils.audiodec.AudioDec`` and the layer classes from /root/reference, loads
synthetic checkpoints written by ``audiodec_b200.synthetic`` (the reference ships no
weights) through the reference's own ``load_transmitter`` / ``load_receiver``
(bin/stream.py:56-77) and dumps inputs + outputs of

    tx_encoder.encode -> tx_encoder.quantize -> rx_encoder.lookup -> decoder.decode
                                                                  (demoFile.py:58-61)

as small ``.npz`` fixtures.  The weights themselves are NOT stored (32 MB); the fixtures
carry the sha256 of the state-dict so tests can assert that the regenerated weights are
the ones used here.  Nothing under tests/ reads /root/reference at test time.
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

from audiodec_b200 import synthetic as S  # noqa: E402

torch.set_num_threads(4)            # demoFile.py:28 default


def expand_buffers(module, b):
    """The reference's pad_buffers are (1,C,P) (layers/conv_layer.py:144-146) so its streaming
    path is batch-1 only; repeating the warmed buffers is the survey-probed way to batch it."""
    from layers.conv_layer import CausalConv1d, CausalConvTranspose1d
    for m in module.modules():
        if isinstance(m, (CausalConv1d, CausalConvTranspose1d)):
            m.pad_buffer = m.pad_buffer.repeat(b, 1, 1)


def load_codec(root, model):
    from utils.audiodec import AudioDec
    sr, enc, dec = S.make_model_zoo(root, model, seed=0)
    os.chdir(root)                      # stats path in the vocoder config is cwd-relative
    a = AudioDec("cpu", "cpu")
    a.load_transmitter(enc)
    a.load_receiver(enc, dec)
    return a


def margins_of(codec, z):
    """relative top-2 margin of every RVQ decision (to classify index mismatches)."""
    layers = codec.tx_encoder.quantizer.codebook.layers
    r = z.transpose(2, 1)
    out = []
    for layer in layers:
        fl = r.reshape(-1, 64)
        dist = fl.pow(2).sum(1, keepdim=True) - 2 * fl @ layer.embed + layer.embed.pow(2).sum(0, keepdim=True)
        top2 = torch.topk(-dist, 2, dim=1).values
        out.append(((top2[:, 0] - top2[:, 1]) / dist.min(1).values.abs()).view(r.shape[:-1]))
        q, _ = layer.forward_index(r)
        r = r - q
    return torch.stack(out)


def run_path(codec, x):
    with torch.no_grad():
        z = codec.tx_encoder.encode(x)
        idx = codec.tx_encoder.quantize(z)
        zq = codec.rx_encoder.lookup(idx)
        y = codec.decoder.decode(zq)
    return z, idx, zq, y


def main():
    scratch = tempfile.mkdtemp(prefix="adec_golden_")
    meta = dict(torch=torch.__version__, threads=4)
    enc_digest = S.state_dict_digest(S.symad_state_dict(seed=0))
    dec_digest = S.state_dict_digest(S.hifigan_state_dict(seed=1))

    # ---- 1. symAD one-shot, 0.25 s @ 48 kHz (demoFile.py path, BASELINE config 1 shortened)
    torch.manual_seed(1337)
    x = 0.1 * torch.randn(1, 1, 12000)
    a = load_codec(scratch, "vctk_sym")
    with torch.no_grad():       # the zq that load_receiver feeds to initial_decoder (bin/stream.py:70,76), from a fresh encoder
        fresh = a._load_encoder(os.path.join(scratch, "exp/autoencoder/symAD_vctk_48000_hop300/checkpoint-200000steps.pkl")).eval()
        zq0 = fresh.initial_encoder(8192, "cpu")
    z, idx, zq, y = run_path(a, x)
    b = load_codec(scratch, "vctk_sym")
    with torch.no_grad():
        zw = b.tx_encoder.encode(x)
        mar = margins_of(b, zw)
    np.savez_compressed(os.path.join(HERE, "symad_oneshot.npz"), x=x.numpy(), z=z.numpy(), idx=idx.numpy(),
                        zq=zq.numpy(), y=y.numpy(), margins=mar.numpy(), warm_zq=zq0.numpy(),
                        enc_digest=enc_digest, **meta)
    print("symad_oneshot: min margin", mar.min().item(), "y absmax", y.abs().max().item())

    # ---- 2. same clip streamed as 8 x 1500-sample chunks (demoStream.py:28 default frame size)
    a = load_codec(scratch, "vctk_sym")
    zs, idxs, ys = [], [], []
    for c in range(8):
        zc, ic, _, yc = run_path(a, x[:, :, c * 1500:(c + 1) * 1500])
        zs.append(zc), idxs.append(ic), ys.append(yc)
    np.savez_compressed(os.path.join(HERE, "symad_stream.npz"), x=x.numpy(), z=torch.cat(zs, -1).numpy(),
                        idx=torch.cat(idxs, -1).numpy(), y=torch.cat(ys, -1).numpy(), chunk=1500,
                        enc_digest=enc_digest, **meta)
    print("symad_stream: idx equal to one-shot:", bool((torch.cat(idxs, -1) == idx).all()),
          "y diff", (torch.cat(ys, -1) - y).abs().max().item())

    # ---- 3. ragged length (T not a multiple of the hop; demoFile.py:61 crops to T)
    torch.manual_seed(7)
    xr = 0.1 * torch.randn(1, 1, 4001)
    a = load_codec(scratch, "vctk_sym")
    z3, idx3, zq3, y3 = run_path(a, xr)
    np.savez_compressed(os.path.join(HERE, "symad_ragged.npz"), x=xr.numpy(), z=z3.numpy(), idx=idx3.numpy(),
                        zq=zq3.numpy(), y=y3.numpy(), enc_digest=enc_digest, **meta)
    print("symad_ragged: frames", z3.shape[-1], "y len", y3.shape[-1])

    # ---- 4. batch of 3 utterances via buffer expansion (reference has no batched streaming call)
    torch.manual_seed(11)
    xb = 0.1 * torch.randn(3, 1, 6000)
    a = load_codec(scratch, "vctk_sym")
    for m in (a.tx_encoder, a.rx_encoder, a.decoder):
        expand_buffers(m, 3)
    with torch.no_grad():
        zb = a.tx_encoder.encode(xb)
        # ResidualVQ.forward_index squeezes dim 1 only when B==1 (vq_module.py:149): (8,3,F) here
        idxb = a.tx_encoder.quantize(zb)
        zqb = torch.sum(torch.nn.functional.embedding(idxb, a.rx_encoder.quantizer.codebook.codebook), dim=0)
        yb = a.decoder.decode(zqb)
    np.savez_compressed(os.path.join(HERE, "symad_batch3.npz"), x=xb.numpy(), z=zb.numpy(), idx=idxb.numpy(),
                        zq=zqb.numpy(), y=yb.numpy(), enc_digest=enc_digest, **meta)
    print("symad_batch3: idx", tuple(idxb.shape), "y", tuple(yb.shape))

    # ---- 5. AD v1 (symAD encoder + HiFi-GAN v1 vocoder), 0.125 s
    torch.manual_seed(1337)
    xv = 0.1 * torch.randn(1, 1, 6000)
    a = load_codec(scratch, "vctk_v1")
    zv, idxv, zqv, yv = run_path(a, xv)
    np.savez_compressed(os.path.join(HERE, "v1_oneshot.npz"), x=xv.numpy(), z=zv.numpy(), idx=idxv.numpy(),
                        zq=zqv.numpy(), y=yv.numpy(), enc_digest=enc_digest, dec_digest=dec_digest, **meta)
    print("v1_oneshot: y absmax", yv.abs().max().item())
    a = load_codec(scratch, "vctk_v1")
    ys = []
    for c in range(4):
        ys.append(run_path(a, xv[:, :, c * 1500:(c + 1) * 1500])[3])
    np.savez_compressed(os.path.join(HERE, "v1_stream.npz"), x=xv.numpy(), y=torch.cat(ys, -1).numpy(), chunk=1500,
                        enc_digest=enc_digest, dec_digest=dec_digest, **meta)

    # ---- 5b. the other released variants (assign_model table, utils/audiodec.py:109-179): one short clip each
    for model, name in (("vctk_v2", "v2_oneshot"), ("vctk_v0", "v0_oneshot"), ("vctk_activate_sym", "aad_oneshot"),
                        ("vctk_c16h320_sym", "c16_oneshot")):
        torch.manual_seed(21)
        xm = 0.1 * torch.randn(1, 1, 6400)
        a = load_codec(scratch, model)
        zm, idxm, zqm, ym = run_path(a, xm)
        a = load_codec(scratch, model)
        ys, idxs = [], []
        for c in range(2):          # and as two 3200-sample chunks (3200 = 10 hops of 320 = not a multiple of 300: ragged for hop 300)
            _, ic, _, yc = run_path(a, xm[:, :, c * 3200:(c + 1) * 3200])
            ys.append(yc), idxs.append(ic)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=xm.numpy(), z=zm.numpy(), idx=idxm.numpy(), zq=zqm.numpy(),
                            y=ym.numpy(), y_chunks=torch.cat(ys, -1).numpy(), idx_chunks=torch.cat(idxs, -1).numpy(), **meta)
        print(name, "idx", tuple(idxm.shape), "y", tuple(ym.shape), "absmax", ym.abs().max().item())

    # ---- 6. layer-level known-answer cases straight from the reference layer classes
    from layers.conv_layer import CausalConv1d, CausalConvTranspose1d
    from layers.vq_module import ResidualVQ
    torch.manual_seed(3)
    cases = {}
    for name, (cin, cout, k, s, d, g, T) in {
        "conv_k7_d3": (8, 12, 7, 1, 3, 1, 50), "conv_k6_s3": (8, 16, 6, 3, 1, 1, 50),
        "conv_k10_s5_ragged": (4, 8, 10, 5, 1, 1, 23), "conv_k11_d5_g3": (12, 12, 11, 1, 5, 3, 40),
        "conv_short_chunk": (8, 8, 7, 1, 9, 1, 5),
    }.items():
        m = CausalConv1d(cin, cout, k, stride=s, dilation=d, groups=g, bias=True)
        with torch.no_grad():
            xs = [torch.randn(1, cin, T), torch.randn(1, cin, T)]
            ysl = [m.inference(xs[0]), m.inference(xs[1])]       # two consecutive chunks: state carry
        cases[name] = dict(w=m.conv.weight.detach().numpy(), b=m.conv.bias.detach().numpy(),
                           x0=xs[0].numpy(), x1=xs[1].numpy(), y0=ysl[0].numpy(), y1=ysl[1].numpy(),
                           cfg=np.array([cin, cout, k, s, d, g, T]))
    for name, (cin, cout, s, T) in {"convtr_s5": (8, 6, 5, 9), "convtr_s3": (4, 4, 3, 1)}.items():
        m = CausalConvTranspose1d(cin, cout, 2 * s, s, bias=True)
        with torch.no_grad():
            xs = [torch.randn(1, cin, T), torch.randn(1, cin, T)]
            ysl = [m.inference(xs[0]), m.inference(xs[1])]
        cases[name] = dict(w=m.deconv.weight.detach().numpy(), b=m.deconv.bias.detach().numpy(),
                           x0=xs[0].numpy(), x1=xs[1].numpy(), y0=ysl[0].numpy(), y1=ysl[1].numpy(),
                           cfg=np.array([cin, cout, 2 * s, s, 1, 1, T]))
    rvq = ResidualVQ(num_quantizers=4, dim=16, codebook_size=32)
    rvq.initial()
    with torch.no_grad():
        xq = torch.randn(1, 25, 16) * 1.5
        zq_, ind = rvq.forward_index(xq, flatten_idx=True)
        lk = rvq.lookup(ind)
    cases["rvq"] = dict(embeds=np.stack([l.embed.numpy() for l in rvq.layers]), x=xq.numpy(), idx=ind.numpy(),
                        zq=zq_.numpy(), lookup=lk.numpy())
    flat = {f"{c}/{k}": v for c, d_ in cases.items() for k, v in d_.items()}
    np.savez_compressed(os.path.join(HERE, "layers.npz"), **flat)
    print("layers:", list(cases))


if __name__ == "__main__":
    main()
