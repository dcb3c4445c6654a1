"""Host-side construction of the mel, chroma and logarithmic filterbanks (setup only, a few KB; the projection runs on
device).
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


def log_filterbank(sr, n_bins=1024, bands_per_octave=24, fmin=30.0, fmax=17000.0, fref=440.0):
    """madmom.audio.filters.LogarithmicFilterbank(bin_frequencies, num_bands=24, fmin, fmax, fref=440, norm_filters=True,
    unique_filters=True) as published (the reference builds it through FilteredSpectrogram, signal.py:57):
    semitone-fraction centre frequencies around A4 -> nearest FFT bins (duplicates dropped) -> overlapping triangular
    filters over consecutive bin triples, each normalised to unit area.  Returns (fb [n_filters, n_bins] float32,
    lo [n_filters], hi [n_filters]) with [lo, hi) = the filter's non-zero bins widened by one neighbour (the range
    ComplexFlux takes its local-group-delay minimum over)."""
    bin_freqs = np.fft.fftfreq(n_bins * 2, 1.0 / sr)[:n_bins]
    left = np.floor(np.log2(float(fmin) / fref) * bands_per_octave)
    right = np.ceil(np.log2(float(fmax) / fref) * bands_per_octave)
    freqs = fref * 2.0 ** (np.arange(left, right) / float(bands_per_octave))
    freqs = freqs[np.searchsorted(freqs, fmin):]
    freqs = freqs[:np.searchsorted(freqs, fmax, "right")]
    idx = np.clip(bin_freqs.searchsorted(freqs), 1, len(bin_freqs) - 1)
    lo_f, hi_f = bin_freqs[idx - 1], bin_freqs[idx]
    idx = np.unique(idx - (freqs - lo_f < hi_f - freqs))
    if len(idx) < 3:
        raise ValueError("log_filterbank: fewer than 3 distinct bins between fmin and fmax")
    rows = []
    for start, center, stop in zip(idx[:-2], idx[1:-1], idx[2:]):
        if stop - start < 2:
            center, stop = start, start + 1
        tri = np.zeros(stop - start)
        tri[:center - start] = np.linspace(0, 1, center - start, endpoint=False)
        tri[center - start:] = np.linspace(1, 0, stop - center, endpoint=False)
        tri /= tri.sum()
        row = np.zeros(n_bins)
        a, b = max(start, 0), min(stop, n_bins)
        row[a:b] = tri[a - start:b - start]
        rows.append(row)
    fb = np.ascontiguousarray(np.stack(rows), dtype=np.float32)
    lo = np.zeros(len(rows), np.int32)
    hi = np.zeros(len(rows), np.int32)
    for i, row in enumerate(fb):
        nz = np.nonzero(row)[0]
        lo[i], hi[i] = max(nz[0] - 1, 0), min(nz[-1] + 2, n_bins)
    return fb, lo, hi
