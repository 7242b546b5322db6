"""Waveform loading (utils/audio.py:7-15).  The reference goes through an old torchaudio API
(``torchaudio.load(path, normalization=True)``: samples scaled to [-1, 1], channels averaged); this reads PCM
WAV files with scipy and applies the same scaling.  The sox-based tempo / gain / noise augmentations of the
reference (utils/audio.py:22-60) shell out to binaries that are not part of this image and are not provided."""
import numpy as np
import scipy.io.wavfile as wavfile


def load_audio(path):
    """-> float32 mono waveform in [-1, 1]."""
    _, data = wavfile.read(path)
    if data.dtype.kind in "iu":
        full = float(1 << (8 * data.dtype.itemsize - 1))
        offset = full if data.dtype.kind == "u" else 0.0
        data = (data.astype(np.float32) - offset) / full
    else:
        data = data.astype(np.float32)
    if data.ndim > 1:
        data = data[:, 0] if data.shape[1] == 1 else data.mean(axis=1)
    return np.ascontiguousarray(data, dtype=np.float32)


def _no_sox(*a, **k):
    raise NotImplementedError("sox-based augmentation (utils/audio.py:22-60) is not available in this build")


get_audio_length = audio_with_sox = augment_audio_with_sox = load_randomly_augmented_audio = _no_sox
