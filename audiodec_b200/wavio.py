"""PCM wav I/O for the file demo (demoFile.py:50-68 uses ``soundfile``; it is not in this image, so fall back to
``scipy.io.wavfile`` with libsndfile's conventions: int16 -> float divides by 32768 on read, float -> PCM_16 multiplies by
32767 and rounds to nearest on write)."""
from __future__ import annotations

import numpy as np


def read_wav(path: str):
    """-> (data float32 (T, C) in [-1, 1), sample_rate)   == ``sf.read(path, always_2d=True)`` (demoFile.py:52)."""
    try:
        import soundfile as sf
        data, fs = sf.read(path, always_2d=True, dtype="float32")
        return data, fs
    except ImportError:
        pass
    from scipy.io import wavfile
    fs, raw = wavfile.read(path)
    if raw.ndim == 1:
        raw = raw[:, None]
    if raw.dtype == np.int16:
        data = raw.astype(np.float32) / 32768.0
    elif raw.dtype == np.int32:
        data = raw.astype(np.float32) / 2147483648.0
    elif raw.dtype == np.uint8:
        data = (raw.astype(np.float32) - 128.0) / 128.0
    else:
        data = raw.astype(np.float32)
    return data, int(fs)


def write_wav_pcm16(path: str, data, sample_rate: int) -> None:
    """``sf.write(path, data, fs, "PCM_16")`` (demoFile.py:63-68); data (T,) or (T, C) float."""
    data = np.asarray(data, dtype=np.float32)
    try:
        import soundfile as sf
        sf.write(path, data, sample_rate, "PCM_16")
        return
    except ImportError:
        pass
    from scipy.io import wavfile
    pcm = np.clip(np.rint(data * 32767.0), -32768, 32767).astype(np.int16)
    wavfile.write(path, int(sample_rate), pcm)
