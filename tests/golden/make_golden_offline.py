#!/usr/bin/env python3
"""Golden vectors of the NON-streaming batch forward, from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, CPU):   python tests/golden/make_golden_offline.py

Drives the reference's base ``Generator`` classes (models/autoencoder/AudioDec.py:23-115, models/vocoder/HiFiGAN.py:24-160)
exactly like codecTest.py:78-95 does: ``encoder.encoder(x) -> encoder.projector -> encoder.quantizer -> decoder.decoder(zq)``
(or ``decoder(zq)`` for HiFi-GAN), on synthetic checkpoints from ``audiodec_b200.synthetic`` (the reference ships none).
Writes tests/golden/offline_*.npz; nothing under tests/ reads /root/reference at test time."""
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

torch.set_num_threads(4)


def codec_test_path(enc, dec, x, vocoder):
    """codecTest.py:78-95 (TestMain.encode / decode) on a (B,1,T) batch."""
    with torch.no_grad():
        h = enc.encoder(x)                       # :84
        z = enc.projector(h)                     # :85
        zq, _, _ = enc.quantizer(z)              # :86
        y = dec(zq) if vocoder else dec.decoder(zq)   # :91-94
    return z, zq, y


def main():
    from models.autoencoder.AudioDec import Generator as GenAD
    from models.vocoder.HiFiGAN import Generator as GenHG
    meta = dict(torch=torch.__version__, threads=4)
    scratch = tempfile.mkdtemp(prefix="adec_golden_off_")
    S.make_model_zoo(scratch, "vctk_v1", seed=0)     # writes stats/synthetic.npy (cwd-relative path in the vocoder params)
    os.chdir(scratch)
    for name, ep, vp, shape, seed in (
        ("offline_symad", S.SYMAD_PARAMS, None, (2, 1, 4801), 31),      # batch of 2, ragged length
        ("offline_aad", S.SYMAAD_PARAMS, None, (1, 1, 3600), 32),
        ("offline_c16", S.SYMAD_C16_PARAMS, None, (1, 1, 3200), 33),
        ("offline_v1", S.SYMAD_PARAMS, S.HIFIGAN_V1_PARAMS, (2, 1, 3000), 34),
        ("offline_v0", S.SYMAD_PARAMS, S.HIFIGAN_V0_PARAMS, (1, 1, 1800), 35),
    ):
        enc = GenAD(**ep)
        enc.load_state_dict(S.symad_state_dict(ep, seed=0))
        enc.eval()
        if vp is None:
            dec = GenAD(**ep)
            dec.load_state_dict(S.symad_state_dict(ep, seed=0))
        else:
            dec = GenHG(**vp)
            dec.load_state_dict(S.hifigan_state_dict(vp, seed=1))
        dec.eval()
        torch.manual_seed(seed)
        x = 0.1 * torch.randn(*shape)
        z, zq, y = codec_test_path(enc, dec, x, vp is not None)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), z=z.numpy(), zq=zq.numpy(), y=y.numpy(), **meta)
        print(name, "z", tuple(z.shape), "y", tuple(y.shape), "y absmax", y.abs().max().item())


if __name__ == "__main__":
    main()
