"""CPU oracle for the audio feature chain (audioreactive/signal.py:31-156 + examples/default.py:6-45).

TEST INFRASTRUCTURE ONLY.  **Parity unpinned** for the library internals: the reference delegates to librosa / madmom
(un-vendored, unpinned in requirements.txt:4-6, absent here), so no golden vector from the reference itself exists for
`onsets`, `chroma`, `rms`.  This file restates (a) the reference's own glue line by line — pinned by
tests/golden/audio_glue.npz for gaussian_filter / percentile_clip / chroma_weight_latents / normalize — and (b) the
published librosa algorithms at the call sites the reference uses, with these documented choices:

  * onsets: `type="rosa"` semantics (`signal.py:50-51`; north_star names librosa): percussive HPSS -> mel power
    spectrogram (128 bands, slaney) -> dB -> lag-1 positive difference -> mean.  The madmom default is a "next" row.
  * STFT: n_fft 2048, hop 512, periodic hann, centred with reflect padding (librosa <= 0.9, the era of the reference).
  * chroma: the CQT front-end of `chroma_cens` is replaced by the STFT chroma filterbank (librosa.filters.chroma,
    tuning fixed to 0) — north_star asks for a cuFFT-fronted chain — followed by the CENS post-processing
    (L1 -> quantise -> 41-frame hann smoothing -> L2) and the cosine k-NN median filter of `raw_chroma` (:130-131).
All maths in float64 unless the reference casts (`.float()` at signal.py:69,95,155).
"""
import numpy as np
import scipy.ndimage
import scipy.signal

N_FFT, HOP = 2048, 512


# ---------------------------------------------------------------------------------------------------------------
# STFT / ISTFT / HPSS  (librosa.core.stft / istft / decompose.hpss / effects.percussive|harmonic)
# ---------------------------------------------------------------------------------------------------------------

def hann(n):
    return scipy.signal.get_window("hann", n, fftbins=True)


def stft(y, n_fft=N_FFT, hop=HOP):
    y = np.asarray(y, np.float64)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return np.fft.rfft(yp[idx] * hann(n_fft)[None, :], axis=1)  # [T, F]


def istft(S, length, n_fft=N_FFT, hop=HOP):
    T = S.shape[0]
    frames = np.fft.irfft(S, n=n_fft, axis=1) * hann(n_fft)[None, :]
    n = n_fft + hop * (T - 1)
    y = np.zeros(n)
    wss = np.zeros(n)
    w2 = hann(n_fft) ** 2
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[t]
        wss[t * hop:t * hop + n_fft] += w2
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2:]
    out = np.zeros(length)
    m = min(length, len(y))
    out[:m] = y[:m]
    return out


def softmask(X, X_ref, power=2.0):
    Z = np.maximum(X, X_ref)
    bad = Z < np.finfo(np.float32).tiny
    Z = np.where(bad, 1.0, Z)
    m = (X / Z) ** power
    r = (X_ref / Z) ** power
    mask = m / (m + r)
    return np.where(bad, 0.0, mask)  # split_zeros=False


def hpss_masks(mag, kernel=31, power=2.0, margin=1.0):
    harm = scipy.ndimage.median_filter(mag, size=(kernel, 1), mode="reflect")   # along time (axis 0 of [T,F])
    perc = scipy.ndimage.median_filter(mag, size=(1, kernel), mode="reflect")   # along frequency
    return softmask(harm, perc * margin, power), softmask(perc, harm * margin, power)


def percussive(y, margin=8.0):
    S = stft(y)
    _, mp = hpss_masks(np.abs(S), margin=margin)
    return istft(S * mp, len(y))


def harmonic(y, margin=16.0):
    S = stft(y)
    mh, _ = hpss_masks(np.abs(S), margin=margin)
    return istft(S * mh, len(y))


# ---------------------------------------------------------------------------------------------------------------
# filterbanks (librosa.filters.mel / chroma)
# ---------------------------------------------------------------------------------------------------------------

def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft=N_FFT, n_mels=128, fmin=0.0, fmax=None):
    fmax = sr / 2 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return w * enorm[:, None]


def chroma_filterbank(sr, n_fft=N_FFT, n_chroma=12, ctroct=5.0, octwidth=2.0):
    freqs = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    a440 = 440.0
    frqbins = n_chroma * np.log2(freqs / (a440 / 16))
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidth = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n2 + 10 * n_chroma, n_chroma) - n2
    w = np.exp(-0.5 * (2 * D / np.tile(binwidth, (n_chroma, 1))) ** 2)
    w = w / np.maximum(np.sqrt((w ** 2).sum(0, keepdims=True)), np.finfo(np.float64).tiny)
    w *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    w = np.roll(w, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(w[:, :1 + n_fft // 2])


# ---------------------------------------------------------------------------------------------------------------
# features
# ---------------------------------------------------------------------------------------------------------------

def power_to_db(S, amin=1e-10, top_db=80.0):
    log_spec = 10.0 * np.log10(np.maximum(amin, S))      # ref = 1.0
    return np.maximum(log_spec, log_spec.max() - top_db)


def onset_strength(y, sr, fmin, fmax):
    S = np.abs(stft(y)) ** 2                               # [T, F]
    mel = S @ mel_filterbank(sr, fmin=fmin, fmax=fmax).T    # [T, 128]
    db = power_to_db(mel)
    env = np.maximum(0.0, db[1:] - db[:-1]).mean(1)
    pad = 1 + N_FFT // (2 * HOP)
    return np.concatenate([np.zeros(pad), env])[:S.shape[0]]


def cens_post(chroma):
    """chroma [T,12] (>=0) -> CENS (librosa.feature.chroma_cens after the CQT front-end)."""
    c = chroma / np.maximum(np.abs(chroma).sum(1, keepdims=True), np.finfo(np.float32).tiny)
    q = np.zeros_like(c)
    for step in (0.4, 0.2, 0.1, 0.05):
        q += (c > step) * 0.25
    win = scipy.signal.get_window("hann", 43, fftbins=False)
    win = win / win.sum()
    sm = scipy.signal.convolve2d(q, win[:, None], mode="same", boundary="fill")
    return sm / np.maximum(np.sqrt((sm ** 2).sum(1, keepdims=True)), np.finfo(np.float32).tiny)


def nn_filter_median_cosine(X):
    """np.minimum(ch, librosa.decompose.nn_filter(ch, aggregate=np.median, metric='cosine')) — signal.py:130-131.
    X [T, C]; neighbours = k nearest frames in cosine distance excluding the frame itself, k = 2*ceil(sqrt(T-1))."""
    T = X.shape[0]
    k = int(min(T - 1, 2 * np.ceil(np.sqrt(T - 2 + 1))))
    nrm = np.sqrt((X ** 2).sum(1))
    Xn = X / np.maximum(nrm, 1e-30)[:, None]
    out = np.empty_like(X)
    for i in range(T):
        d = 1.0 - Xn @ Xn[i]
        d[i] = np.inf
        nb = np.argpartition(d, k - 1)[:k]
        out[i] = np.median(X[nb], axis=0)
    return np.minimum(X, out)


def resample(x, num):
    """scipy.signal.resample along axis 0 (Fourier method)."""
    return scipy.signal.resample(np.asarray(x, np.float64), num, axis=0)


def gaussian_filter(x, sigma, causal=None, smf=1.0):
    """audioreactive/signal.py:319-368 on a [T, ...] float32 array (circular along time)."""
    x = np.asarray(x, np.float32)
    T = x.shape[0]
    flat = x.reshape(T, -1)
    radius = min(int(sigma * 4 * smf), 3 * T)
    k = np.arange(-radius, radius + 1, dtype=np.float32)
    g = np.exp(np.float32(-0.5 / sigma ** 2) * k ** 2).astype(np.float32)
    if causal is not None:
        g[radius + 1:] *= 0 if not isinstance(causal, float) else np.float32(causal)
    g = (g / g.sum()).astype(np.float32)
    if radius > T:
        xp = np.concatenate([flat, flat, flat], 0)
        z = np.zeros((radius - T, flat.shape[1]), np.float32)
        xp = np.concatenate([z, xp, z], 0)
    else:
        xp = np.concatenate([flat[T - radius:], flat, flat[:radius]], 0) if radius > 0 else flat
    out = np.zeros_like(flat, dtype=np.float64)
    for j in range(2 * radius + 1):
        out += g[j].astype(np.float64) * xp[j:j + T]
    return out.astype(np.float32).reshape(x.shape)


def percentile_clip(sig, p):
    """signal.py:273-292: clamp to the p-th percentile of the strict local maxima, then /max."""
    s = np.asarray(sig, np.float32)
    n = len(s)
    idx = np.arange(n)
    plus = s[np.clip(idx + 1, 0, n - 1)]
    minus = s[np.clip(idx - 1, 0, n - 1)]
    peaks = s[(s > plus) & (s > minus)]
    k = 1 + round(0.01 * float(p) * (len(peaks) - 1))
    thr = np.sort(peaks)[k - 1]
    s = np.clip(s, 0, thr)
    return s / s.max()


# ---- madmom (un-vendored, unpinned in requirements.txt; absent here): restated from its published source -------------
def mm_log_filterbank(sr, n_bins=1024, bands_per_octave=24, fmin=30.0, fmax=17000.0, fref=440.0):
    """madmom.audio.filters: log_frequencies -> frequencies2bins(unique) -> TriangularFilter.filters(norm, overlap)
    -> Filterbank.from_filters.  Returns [n_bins, n_filters]."""
    bin_freqs = np.fft.fftfreq(n_bins * 2, 1.0 / sr)[:n_bins]
    left = np.floor(np.log2(float(fmin) / fref) * bands_per_octave)
    right = np.ceil(np.log2(float(fmax) / fref) * bands_per_octave)
    f = fref * 2.0 ** (np.arange(left, right) / float(bands_per_octave))
    f = f[np.searchsorted(f, fmin):]
    f = f[:np.searchsorted(f, fmax, "right")]
    idx = np.clip(bin_freqs.searchsorted(f), 1, len(bin_freqs) - 1)
    idx = idx - ((f - bin_freqs[idx - 1]) < (bin_freqs[idx] - f))
    bins = np.unique(idx)
    filters = []
    i = 0
    while i + 3 <= len(bins):
        start, center, stop = bins[i:i + 3]
        i += 1
        if stop - start < 2:
            center, stop = start, start + 1
        data = np.zeros(stop - start)
        data[:center - start] = np.linspace(0, 1, center - start, endpoint=False)
        data[center - start:] = np.linspace(1, 0, stop - center, endpoint=False)
        filters.append((start, data / data.sum()))
    fb = np.zeros((n_bins, len(filters)), np.float32)
    for b, (start, data) in enumerate(filters):
        seg = fb[start:start + len(data), b]
        np.maximum(data[:len(seg)], seg, out=seg)
    return fb


def mm_stft(y, frame_size=2048, hop=441):
    """FramedSignal(origin 0, end 'normal') + ShortTimeFourierTransform(window=np.hanning, circular_shift=True):
    complex spectrogram [T, frame_size/2] (no Nyquist bin)."""
    y = np.asarray(y, np.float32)
    T = int(np.ceil(len(y) / float(hop)))
    win = np.hanning(frame_size).astype(np.float32)
    out = np.zeros((T, frame_size // 2), np.complex64)
    pad = np.concatenate([np.zeros(frame_size // 2, np.float32), y, np.zeros(frame_size + hop, np.float32)])
    for t in range(T):
        fr = pad[t * hop:t * hop + frame_size] * win           # starts at t*hop - frame_size/2 in the unpadded signal
        fr = np.concatenate([fr[frame_size // 2:], fr[:frame_size // 2]])
        out[t] = np.fft.fft(fr)[:frame_size // 2]
    return out


def onset_strength_mm(y, sr, fmin, fmax, frame_size=2048, hop=441):
    """signal.py:53-66: sum of madmom.features.onsets.{spectral_diff, spectral_flux, superflux, complex_flux,
    modified_kullback_leibler} on the 24-bands/octave filtered magnitude spectrogram."""
    from scipy.ndimage import maximum_filter

    S = mm_stft(y, frame_size, hop)
    fb = mm_log_filterbank(sr, frame_size // 2, 24, fmin, fmax)
    spec = np.abs(S).astype(np.float32) @ fb
    win = np.hanning(frame_size)
    df = int(max(1, round((len(win) / 2 - np.argmax(win > 0.5 * max(win))) / hop)))

    def diff(max_bins=None):
        ref = maximum_filter(spec, size=(1, max_bins)) if max_bins else spec
        d = np.zeros_like(spec)
        d[df:] = spec[df:] - ref[:-df]
        return np.maximum(d, 0)

    sd = np.sum(diff() ** 2, axis=1)
    sf = np.sum(diff(), axis=1)
    su = np.sum(diff(3), axis=1)
    phase = np.angle(S)
    up = np.unwrap(phase)
    up[:, :-1] -= up[:, 1:]
    up[:, -1] = 0
    lgd = maximum_filter(np.abs(up) / np.pi, size=[3, 1])
    mask = np.zeros_like(spec)
    for b in range(spec.shape[1]):
        nz = np.nonzero(fb[:, b])[0]
        mask[:, b] = np.amin(lgd[:, max(nz[0] - 1, 0):min(nz[-1] + 2, lgd.shape[1])], axis=1)
    cf = np.sum(diff(3) * mask, axis=1)
    mkl = np.zeros_like(spec)
    mkl[1:] = spec[1:] / (spec[:-1] + np.float32(np.spacing(1)))
    kl = np.mean(np.log(1 + mkl), axis=1)
    return (sd + sf + su + cf + kl).astype(np.float32)


def onsets(audio, sr, n_frames, margin=8, fmin=20, fmax=8000, smooth=1, clip=100, power=1, smf=1.0, type="rosa"):
    """signal.py:31-73."""
    y_perc = percussive(audio, margin=margin)
    o = onset_strength_mm(y_perc, sr, fmin, fmax) if type == "mm" else onset_strength(y_perc, sr, fmin, fmax)
    o = np.clip(resample(o, n_frames), o.min(), o.max()).astype(np.float32)
    o = gaussian_filter(o, smooth, causal=0, smf=smf)
    o = percentile_clip(o, clip)
    return o ** power


def chroma(audio, sr, n_frames, margin=16, notes=12):
    """signal.py:136-156 with the STFT chroma front-end (see module docstring)."""
    y_harm = harmonic(audio, margin=margin)
    S = np.abs(stft(y_harm)) ** 2
    raw = S @ chroma_filterbank(sr).T
    ch = cens_post(raw)
    ch = nn_filter_median_cosine(ch)
    ch = resample(ch, n_frames)
    keep = np.argsort(np.median(ch, axis=0))[:notes]
    ch = ch[:, keep]
    return (ch / ch.sum(1)[:, None]).astype(np.float32)


def rms(y, sr, n_frames, fmin=20, fmax=8000, smooth=180, clip=50, power=6, smf=1.0):
    """signal.py:76-99."""
    sos = scipy.signal.butter(12, [fmin, fmax], "bp", fs=sr, output="sos")
    y_filt = scipy.signal.sosfilt(sos, np.asarray(y, np.float64))
    S = np.abs(stft(y_filt))
    r = np.sqrt((2 * (S ** 2).sum(1) - S[:, 0] ** 2 - S[:, -1] ** 2) / N_FFT ** 2)     # librosa.feature.rms(S=...)
    r = np.clip(resample(r, n_frames), r.min(), r.max()).astype(np.float32)
    r = gaussian_filter(r, smooth, causal=0.05, smf=smf)
    r = percentile_clip(r, clip)
    return r ** power


def chroma_weight_latents(chroma_t, latents):
    """latent.py:15-26"""
    return (chroma_t[..., None, None] * latents[None, ...]).sum(1)


def default_get_latents(selection, chroma_t, lo, hi, smf=1.0):
    """examples/default.py:12-25"""
    lat = gaussian_filter(chroma_weight_latents(chroma_t, selection).astype(np.float32), 4, smf=smf)
    lo_, hi_ = lo[:, None, None], hi[:, None, None]
    lat = hi_ * selection[[-4]] + (1 - hi_) * lat
    lat = lo_ * selection[[-7]] + (1 - lo_) * lat
    return gaussian_filter(lat.astype(np.float32), 2, causal=0.2, smf=smf)
