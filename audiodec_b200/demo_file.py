"""File demo with the reference's command line (demoFile.py:20-70):

    python -m audiodec_b200.demo_file --model vctk_v1 -i input.wav -o output.wav [--cuda 0]

wav -> AudioDec.load_transmitter / load_receiver -> encode -> quantize -> lookup -> decode -> PCM_16 wav, on the GPU.
``--cuda -1`` (the reference's CPU mode) is refused: this implementation has no CPU path."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from audiodec_b200.utils.audiodec import AudioDec, assign_model
from audiodec_b200.wavio import read_wav, write_wav_pcm16


def run_file(audiodec, data: np.ndarray) -> np.ndarray:
    """(T, C) float -> (T, C) float through the four calls of demoFile.py:55-62 (channels ride the batch dimension)."""
    x = torch.tensor(np.expand_dims(data.transpose(1, 0), axis=1), dtype=torch.float).to(audiodec.tx_device)   # (T,C) -> (C,1,T)
    with torch.no_grad():
        z = audiodec.tx_encoder.encode(x)
        idx = audiodec.tx_encoder.quantize(z)
        zq = audiodec.rx_encoder.lookup(idx)
        y = audiodec.decoder.decode(zq)[:, :, :x.size(-1)]
    return y.squeeze(1).transpose(1, 0).cpu().numpy()


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--model", type=str, default="libritts_v1")
    parser.add_argument("-i", "--input", type=str, required=True)
    parser.add_argument("-o", "--output", type=str, required=True)
    parser.add_argument("--cuda", type=int, default=0)
    parser.add_argument("--num_threads", type=int, default=4)
    args = parser.parse_args(argv)
    if args.cuda < 0:
        raise SystemExit("audiodec_b200 has no CPU path: pass --cuda <ordinal>")
    device = f"cuda:{args.cuda}"
    torch.set_num_threads(args.num_threads)
    sample_rate, encoder_checkpoint, decoder_checkpoint = assign_model(args.model)
    if not os.path.exists(args.input):
        raise ValueError(f"Input file {args.input} does not exist!")
    print("AudioDec initinalizing!")
    audiodec = AudioDec(tx_device=device, rx_device=device)
    audiodec.load_transmitter(encoder_checkpoint)
    audiodec.load_receiver(encoder_checkpoint, decoder_checkpoint)
    data, fs = read_wav(args.input)
    assert fs == sample_rate, f"data ({fs}Hz) is not matched to model ({sample_rate}Hz)!"
    print("Encode/Decode...")
    write_wav_pcm16(args.output, run_file(audiodec, data), fs)
    print(f"Output {args.output}!")


if __name__ == "__main__":
    main()
