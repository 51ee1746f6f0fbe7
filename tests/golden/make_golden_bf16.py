#!/usr/bin/env python3
"""Golden vectors for the vocoder's bf16 mode (BASELINE configs[2]), from the UNMODIFIED reference on CPU:

    python tests/golden/make_golden_bf16.py          (build container only: needs /root/reference)

AudioDec v1 = symAD encoder (fp32) + HiFi-GAN v1 vocoder.  The fixture holds, for one 0.25 s clip through the reference's own
load_transmitter / load_receiver / encode / quantize / lookup (all fp32), the decoder output
  y_fp32 : decoder.decode(zq) in fp32 (the 1e-4 oracle of every other test), and
  y_bf16 : the same decoder after `decoder.to(torch.bfloat16)` fed zq.bfloat16() - what the reference produces when asked for bf16
           (weights, stats, pad_buffers and every activation in bf16; torch CPU bf16 convolutions).
The reference defines no tolerance for reduced precision; the test derives one from these two.
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
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

from audiodec_b200 import synthetic as S  # noqa: E402
from make_golden import load_codec  # noqa: E402

torch.set_num_threads(4)


def main():
    scratch = tempfile.mkdtemp(prefix="adec_golden_bf16_")
    torch.manual_seed(1337)
    x = 0.1 * torch.randn(1, 1, 12000)
    a = load_codec(scratch, "vctk_v1")
    with torch.no_grad():
        z = a.tx_encoder.encode(x)
        idx = a.tx_encoder.quantize(z)
        zq = a.rx_encoder.lookup(idx)
        y32 = a.decoder.decode(zq)
    b = load_codec(scratch, "vctk_v1")            # fresh warm state, then cast the whole vocoder (pad_buffers included)
    b.decoder.to(torch.bfloat16)
    with torch.no_grad():
        y16 = b.decoder.decode(zq.to(torch.bfloat16)).float()
    err = (y16 - y32).abs()
    snr = 10 * torch.log10(y32.pow(2).mean() / (y16 - y32).pow(2).mean())
    print(f"reference bf16 vs fp32: max abs {err.max().item():.4e}, rms {err.pow(2).mean().sqrt().item():.4e}, SNR {snr.item():.1f} dB, "
          f"y rms {y32.pow(2).mean().sqrt().item():.3f}")
    np.savez_compressed(os.path.join(HERE, "v1_bf16.npz"), x=x.numpy(), idx=idx.numpy(), zq=zq.numpy(), y_fp32=y32.numpy(), y_bf16=y16.numpy(),
                        enc_digest=S.state_dict_digest(S.symad_state_dict(seed=0)), dec_digest=S.state_dict_digest(S.hifigan_state_dict(seed=1)),
                        torch=torch.__version__, threads=4)


if __name__ == "__main__":
    main()
