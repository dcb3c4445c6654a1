"""Host-side construction of the mel and chroma filterbanks (setup only, a few KB; the projection runs on device).
Slaney-style mel bank and the STFT chroma bank as published for librosa.filters.mel / librosa.filters.chroma
(the reference calls them through `rosa.onset.onset_strength`, signal.py:51, and `rosa.feature.chroma_*`, :115-119)."""
import numpy as np


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel(sr, n_fft=2048, n_mels=128, fmin=0.0, fmax=None):
    fmax = sr / 2 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return np.ascontiguousarray(w, dtype=np.float32)


def chroma(sr, n_fft=2048, n_chroma=12, ctroct=5.0, octwidth=2.0):
    freqs = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    frqbins = n_chroma * np.log2(freqs / (440.0 / 16))
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidth = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n2 + 10 * n_chroma, n_chroma) - n2
    w = np.exp(-0.5 * (2 * D / np.tile(binwidth, (n_chroma, 1))) ** 2)
    w = w / np.maximum(np.sqrt((w ** 2).sum(0, keepdims=True)), np.finfo(np.float64).tiny)
    w *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    w = np.roll(w, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(w[:, :1 + n_fft // 2], dtype=np.float32)
