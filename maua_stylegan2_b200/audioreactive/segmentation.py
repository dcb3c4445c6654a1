"""`laplacian_segmentation` of the reference (audioreactive/signal.py:159-240: the librosa "Laplacian segmentation" recipe
of McFee & Ellis 2014) on the device path.

The sample-rate work runs on the GPU through libmaua_b200.so — STFT (cuFFT), the constant-Q log-frequency projection, the
mel projection for the MFCCs and the onset envelope that drives the beat tracker; what is left is O(n_beats^2) on a few
hundred beats (beat-synchronous aggregation, recurrence graph, time-lag median filter, eigendecomposition of the
normalised Laplacian, k-means) and stays on the host with numpy / scipy / scikit-learn, exactly the libraries the
reference itself uses for those steps.

librosa is un-vendored and absent: its functions are restated from their published algorithms and defaults ("parity
unpinned") —
  * `librosa.cqt` (36 bins / octave x 7 octaves from C1) -> a constant-Q triangular projection of the 2048-point power
    STFT (north_star: cuFFT-fronted); below ~130 Hz the STFT bins are wider than the CQT bins, so the lowest two octaves
    are smoother than a true multi-rate CQT;
  * `librosa.beat.beat_track` -> global tempo from the autocorrelation of the onset envelope with a log-normal prior
    around 120 BPM, then the dynamic-programming tracker of Ellis (2007) with tightness 100, trim=False;
  * `librosa.segment.recurrence_matrix(width=3, mode="affinity", sym=True)`, `timelag_filter`, `librosa.feature.mfcc`
    (128 mels, dB, orthonormal DCT-II, 20 coefficients), `librosa.util.sync` — restated line by line below.
K-means is seeded (the reference's is not): the same audio always gives the same sections.
"""
import numpy as np
import scipy.fft
import scipy.linalg
import scipy.ndimage
import scipy.sparse.csgraph
import torch as th

from . import filters
from . import signal as S

BINS_PER_OCTAVE = 12 * 3
N_OCTAVES = 7
FMIN_C1 = 32.70319566257483


def cq_filterbank(sr, n_fft=S.N_FFT, n_bins=BINS_PER_OCTAVE * N_OCTAVES, bins_per_octave=BINS_PER_OCTAVE, fmin=FMIN_C1):
    """[n_bins, n_fft/2+1] triangular constant-Q weights on the STFT bins (centres fmin * 2^(k/bpo), unit-sum rows)."""
    freqs = np.arange(n_fft // 2 + 1) * sr / n_fft
    centres = fmin * 2.0 ** (np.arange(-1, n_bins + 1) / bins_per_octave)
    fb = np.zeros((n_bins, len(freqs)), np.float32)
    for k in range(n_bins):
        lo, c, hi = centres[k], centres[k + 1], centres[k + 2]
        lo, hi = min(lo, c - sr / n_fft), max(hi, c + sr / n_fft)     # never narrower than one STFT bin
        w = np.minimum((freqs - lo) / (c - lo), (hi - freqs) / (hi - c))
        w = np.maximum(w, 0.0)
        if w.sum() > 0:
            fb[k] = w / w.sum()
    return fb


def _db(x, ref, amin, top_db=80.0):
    out = 10.0 * np.log10(np.maximum(x, amin)) - 10.0 * np.log10(max(ref, amin))
    return np.maximum(out, out.max() - top_db)


def device_features(signal, sr):
    """(C [n_cq, T] dB relative to the maximum, mfcc [20, T], onset envelope [T]) — STFT and projections on the GPU."""
    y = S._to_dev(signal)
    spec = S.stft(y)                                                       # [T, 1025, 2]
    cq_power = S._filterbank(spec, cq_filterbank(sr)).cpu().numpy().T      # sum |S|^2 * w  ->  [n_cq, T]
    C = _db(cq_power, cq_power.max(), amin=1e-10)                          # amplitude_to_db(|C|, ref=max) = power dB
    mel = S._filterbank(spec, filters.mel(sr, S.N_FFT, 128, 0.0, None)).cpu().numpy().T
    mfcc = scipy.fft.dct(_db(mel, 1.0, amin=1e-10), axis=0, type=2, norm="ortho")[:20]
    onset = S.onset_strength(y, sr, 0.0, None).cpu().numpy().astype(np.float64)
    return C, mfcc, onset


def estimate_tempo(onset, sr, hop=S.HOP, start_bpm=120.0, std_bpm=1.0, max_tempo=320.0):
    """librosa.beat.tempo (aggregate over the whole clip): autocorrelation x log-normal prior."""
    fps = sr / hop
    n = len(onset)
    ac_size = int(min(n, round(8.0 * fps)))
    x = onset - onset.mean()
    spec = np.fft.rfft(x, 2 * n)
    ac = np.fft.irfft(spec * np.conj(spec))[:ac_size]
    ac = ac / (ac[0] + 1e-12)
    bpms = np.zeros(ac_size)
    bpms[1:] = 60.0 * fps / np.arange(1, ac_size)
    prior = np.zeros(ac_size)
    prior[1:] = np.exp(-0.5 * ((np.log2(bpms[1:]) - np.log2(start_bpm)) / std_bpm) ** 2)
    prior[bpms > max_tempo] = 0
    best = int(np.argmax(ac * prior))
    return bpms[best] if best > 0 else start_bpm


def beat_track(onset, sr, hop=S.HOP, tightness=100.0):
    """Ellis (2007) dynamic-programming beat tracker as in librosa.beat.beat_track(trim=False): (tempo, beat frames)."""
    fps = sr / hop
    bpm = estimate_tempo(onset, sr, hop)
    period = int(round(60.0 * fps / bpm))
    if period < 1 or onset.std() == 0:
        return bpm, np.zeros(0, int)
    o = onset / onset.std()
    win = np.exp(-0.5 * (np.arange(-period, period + 1) * 32.0 / period) ** 2)
    local = scipy.signal.convolve(o, win, "same")
    n = len(local)
    backlink = np.full(n, -1, int)
    cumscore = np.zeros(n)
    window = np.arange(-2 * period, -int(round(period / 2)) + 1)
    txwt = -tightness * np.log(-window / period) ** 2
    first = True
    for i in range(n):
        idx = i + window
        valid = idx >= 0
        score = np.full(len(window), -np.inf)
        score[valid] = txwt[valid] + cumscore[idx[valid]]
        best = int(np.argmax(score))
        cumscore[i] = local[i] + (score[best] if np.isfinite(score[best]) else 0.0)
        if first and local[i] < 0.01 * local.max():
            backlink[i] = -1
        else:
            backlink[i] = idx[best] if np.isfinite(score[best]) else -1
            first = False
    # last beat: the best local maximum of the cumulative score above half the median of the maxima
    maxes = np.flatnonzero((cumscore[1:-1] > cumscore[:-2]) & (cumscore[1:-1] >= cumscore[2:])) + 1
    if len(maxes) == 0:
        return bpm, np.zeros(0, int)
    thresh = 0.5 * np.median(cumscore[maxes])
    tail = int(maxes[cumscore[maxes] > thresh].max())
    beats = [tail]
    while backlink[beats[-1]] >= 0:
        beats.append(int(backlink[beats[-1]]))
    return bpm, np.array(beats[::-1], int)


def sync(data, frames, aggregate=np.mean):
    """librosa.util.sync(data, idx, aggregate, pad=True) along the last axis."""
    n = data.shape[-1]
    bounds = np.unique(np.concatenate([[0], np.clip(frames, 0, n), [n]]))
    return np.stack([aggregate(data[..., a:b], axis=-1) for a, b in zip(bounds[:-1], bounds[1:]) if b > a], axis=-1)


def recurrence_affinity(X, width=3):
    """librosa.segment.recurrence_matrix(X, width=3, mode='affinity', sym=True) for X [d, n]."""
    n = X.shape[1]
    k = int(min(n - 1, max(1, 2 * np.ceil(np.sqrt(max(n - 2 * width + 1, 1))))))
    D = np.sqrt(np.maximum(((X[:, :, None] - X[:, None, :]) ** 2).sum(0), 0.0))
    banned = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) < width
    Dm = np.where(banned, np.inf, D)
    order = np.argsort(Dm, axis=1)[:, :k]
    knn = np.zeros((n, n), bool)
    rows = np.repeat(np.arange(n), order.shape[1])
    keep = np.isfinite(Dm[rows, order.ravel()])
    knn[rows[keep], order.ravel()[keep]] = True
    knn = knn & knn.T                                      # sym=True: mutual neighbours
    kth = np.array([np.sort(Dm[i][np.isfinite(Dm[i])])[min(k, np.isfinite(Dm[i]).sum()) - 1] if np.isfinite(Dm[i]).any()
                    else 1.0 for i in range(n)])
    bandwidth = max(float(np.median(kth)), 1e-12)
    return np.where(knn, np.exp(-D / bandwidth), 0.0)


def timelag_median(R, size=7):
    """librosa.segment.timelag_filter(scipy.ndimage.median_filter)(R, size=(1, size)): filter along time in lag space."""
    n = R.shape[0]
    lag = np.zeros((2 * n, n), R.dtype)                      # recurrence_to_lag(pad=True): column j rolled down by j
    for j in range(n):
        lag[:, j] = np.roll(np.concatenate([R[:, j], np.zeros(n, R.dtype)]), j)
    lag = scipy.ndimage.median_filter(lag, size=(1, size), mode="mirror")
    out = np.zeros_like(R)
    for j in range(n):
        out[:, j] = np.roll(lag[:, j], -j)[:n]
    return out


def laplacian_segmentation(signal, sr, k=5, plot=False):
    """Segments the audio with pattern recurrence analysis (audioreactive/signal.py:159-240).
    Returns (list of segment start times in seconds, list of segment labels)."""
    import sklearn.cluster

    C, mfcc, onset = device_features(signal, sr)
    tempo, beats = beat_track(onset, sr)
    if len(beats) < max(2 * k, 8):                            # too short / no pulse: fall back to a regular one-second grid
        beats = np.arange(0, C.shape[1], max(int(round(sr / S.HOP)), 1))
    Csync = sync(C, beats, aggregate=np.median)
    R = recurrence_affinity(Csync, width=3)
    Rf = timelag_median(R, size=7)
    Msync = sync(mfcc, beats)
    path_distance = np.sum(np.diff(Msync, axis=1) ** 2, axis=0)
    sigma = max(float(np.median(path_distance)), 1e-12)
    path_sim = np.exp(-path_distance / sigma)
    R_path = np.diag(path_sim, k=1) + np.diag(path_sim, k=-1)
    deg_path, deg_rec = R_path.sum(1), Rf.sum(1)
    mu = deg_path.dot(deg_path + deg_rec) / max(np.sum((deg_path + deg_rec) ** 2), 1e-12)
    A = mu * Rf + (1 - mu) * R_path
    Lap = scipy.sparse.csgraph.laplacian(A, normed=True)
    _, evecs = scipy.linalg.eigh(Lap)
    evecs = scipy.ndimage.median_filter(evecs, size=(9, 1))
    Cnorm = np.cumsum(evecs ** 2, axis=1) ** 0.5
    k = int(min(k, evecs.shape[1]))
    X = evecs[:, :k] / np.maximum(Cnorm[:, k - 1:k], 1e-12)
    seg_ids = sklearn.cluster.KMeans(n_clusters=k, n_init=10, random_state=0).fit_predict(X)

    bound_beats = 1 + np.flatnonzero(seg_ids[:-1] != seg_ids[1:])
    bound_beats = np.unique(np.concatenate([[0], bound_beats]))              # librosa.util.fix_frames(x_min=0)
    bound_segs = [int(s) for s in seg_ids[bound_beats]]
    seg_starts = np.unique(np.concatenate([[0], np.clip(beats, 0, C.shape[1]), [C.shape[1]]]))[:-1]   # frame of each synced column
    bound_frames = np.minimum(seg_starts[np.minimum(bound_beats, len(seg_starts) - 1)], C.shape[1] - 1)
    bound_times = bound_frames * S.HOP / 22050.0              # librosa.frames_to_time defaults (sr=22050, hop 512), as called
    bound_times[0] = 0.0
    if plot:
        raise NotImplementedError("plot=True needs matplotlib / librosa.display, which the device path does not ship")
    return [float(t) for t in bound_times], bound_segs
