"""TEST INFRASTRUCTURE ONLY - numpy statement of the index bitstream.

The reference defines no wire format: `AudioCodecStreamer` puts the int64 `(Nq,F)` index tensor itself on a queue
(bin/stream.py:224), and `quantize` emits flat indices `j + codebook_size*i` (layers/vq_module.py:145-147).  The format
stated here is therefore this repo's own; parity = the CUDA pack kernel reproduces these bytes exactly and
unpack(pack(idx)) == idx for every index tensor the quantiser can emit.  Nothing under audiodec_b200/ imports this file.

Frame = Nq local indices (flat - i*N) of bits = ceil(log2 N) each, stage 0 first, little-endian bit order
(value k occupies stream bits [k*bits, (k+1)*bits), bit b of the stream is bit b%8 of byte b//8), zero-padded to whole bytes.
"""
import numpy as np


def index_bits(codebook_size: int) -> int:
    b = 1
    while (1 << b) < codebook_size:
        b += 1
    return b


def frame_bytes(codebook_num: int, codebook_size: int) -> int:
    return (codebook_num * index_bits(codebook_size) + 7) // 8


def pack_indices(idx: np.ndarray, codebook_size: int) -> np.ndarray:
    """idx (Nq,B,F) int64 flat -> (B,F,frame_bytes) uint8.  Plain loops over bits: small cases only."""
    nq, B, F = idx.shape
    bits = index_bits(codebook_size)
    out = np.zeros((B, F, frame_bytes(nq, codebook_size)), np.uint8)
    for b in range(B):
        for f in range(F):
            pos = 0
            for i in range(nq):
                v = int(idx[i, b, f]) - i * codebook_size
                assert 0 <= v < codebook_size, "index out of range"
                for k in range(bits):
                    if (v >> k) & 1:
                        out[b, f, pos >> 3] |= 1 << (pos & 7)
                    pos += 1
    return out


def unpack_indices(packed: np.ndarray, codebook_num: int, codebook_size: int) -> np.ndarray:
    """(B,F,frame_bytes) uint8 -> idx (Nq,B,F) int64 flat."""
    B, F, nb = packed.shape
    bits = index_bits(codebook_size)
    assert nb == frame_bytes(codebook_num, codebook_size)
    idx = np.zeros((codebook_num, B, F), np.int64)
    for b in range(B):
        for f in range(F):
            pos = 0
            for i in range(codebook_num):
                v = 0
                for k in range(bits):
                    v |= ((int(packed[b, f, pos >> 3]) >> (pos & 7)) & 1) << k
                    pos += 1
                idx[i, b, f] = v + i * codebook_size
    return idx
