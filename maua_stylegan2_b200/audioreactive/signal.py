"""Device-side drop-in for the signal half of the reference's `audioreactive/signal.py`.

Same function names and arguments (`onsets`, `rms`, `chroma`, `raw_chroma`, `gaussian_filter`, `percentile`,
`percentile_clip`, `normalize`, `compress`, `expand`, `set_SMF`, `load_audio`), but the work runs on the GPU through
libmaua_b200.so (csrc/audio.cu): cuFFT STFT -> HPSS -> ISTFT -> STFT -> mel/onset | chroma | rms -> Fourier resample ->
gaussian filter -> percentile clip.  Results are CUDA tensors (the reference returns CPU tensors; set
`OUTPUT_DEVICE = "cpu"` for legacy hook files that mix them with CPU latents).

Differences from the reference, all forced by its un-vendored dependencies (parity unpinned):
  * `onsets(type="mm")` (madmom, the reference default) and `type="rosa"` (librosa) are both restated from the
    libraries' published algorithms.
  * `chroma`: the CQT front-end of chroma_cens is replaced by the STFT chroma filterbank (north_star: cuFFT-fronted);
    CENS post-processing and the cosine k-NN median filter follow the published librosa algorithms.
  * `laplacian_segmentation` lives in segmentation.py (device front-end, host spectral clustering).
"""
import math
import os
from pathlib import Path

import numpy as np
import torch as th

from .. import _lib as L
from . import filters

SMF = 1  # smoothing multiplier, set by generate() from the rendering fps (signal.py:18-23)
OUTPUT_DEVICE = "cuda"
N_FFT, HOP = 2048, 512


def set_SMF(smf):
    global SMF
    SMF = smf


def _dev():
    return th.device("cuda", th.cuda.current_device())


def _to_dev(x):
    if isinstance(x, np.ndarray):
        x = th.from_numpy(np.ascontiguousarray(x))
    return x.to(device=_dev(), dtype=th.float32).contiguous()


def _out(t):
    return t if OUTPUT_DEVICE == "cuda" else t.to(OUTPUT_DEVICE)


def _n_stft_frames(n):
    return 1 + n // HOP  # centred STFT: 1 + (n + n_fft - n_fft) // hop


# ---------------------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------------------

def stft(y):
    y = _to_dev(y)
    T = _n_stft_frames(y.numel())
    spec = th.empty((T, N_FFT // 2 + 1, 2), device=y.device, dtype=th.float32)
    ws = th.empty(T * N_FFT, device=y.device, dtype=th.float32)
    L.call("maua_audio_stft_f32", y.data_ptr(), y.numel(), spec.data_ptr(), ws.data_ptr(), N_FFT, HOP, T,
           L.stream_ptr(y.device))
    return spec


def istft(spec, length):
    T = spec.shape[0]
    y = th.empty(length, device=spec.device, dtype=th.float32)
    ws = th.empty(T * N_FFT, device=spec.device, dtype=th.float32)
    L.call("maua_audio_istft_f32", spec.data_ptr(), y.data_ptr(), length, ws.data_ptr(), N_FFT, HOP, T,
           L.stream_ptr(spec.device))
    return y


def _hpss_component(y, margin, which):
    """librosa.effects.harmonic (which=0) / percussive (which=1): STFT -> soft-masked median HPSS -> ISTFT."""
    y = _to_dev(y)
    spec = stft(y)
    out = th.empty_like(spec)
    mag = th.empty(spec.shape[:2], device=y.device, dtype=th.float32)
    L.call("maua_audio_hpss_f32", spec.data_ptr(), out.data_ptr(), mag.data_ptr(), spec.shape[0], spec.shape[1],
           float(margin), 2.0, which, L.stream_ptr(y.device))
    return istft(out, y.numel())


def harmonic(y, margin=16):
    return _hpss_component(y, margin, 0)


def percussive(y, margin=8):
    return _hpss_component(y, margin, 1)


def _filterbank(spec, fb):
    fb = th.from_numpy(fb).to(spec.device)
    out = th.empty((spec.shape[0], fb.shape[0]), device=spec.device, dtype=th.float32)
    L.call("maua_audio_filterbank_f32", spec.data_ptr(), fb.data_ptr(), out.data_ptr(), spec.shape[0], spec.shape[1],
           fb.shape[0], L.stream_ptr(spec.device))
    return out


def onset_strength(y, sr, fmin, fmax):
    """librosa.onset.onset_strength(y, sr, fmin=, fmax=): 128-band mel power -> dB -> lag-1 positive diff -> mean."""
    spec = stft(y)
    mel = _filterbank(spec, filters.mel(sr, N_FFT, 128, fmin, fmax))
    env = th.empty(spec.shape[0], device=spec.device, dtype=th.float32)
    scalar = th.empty(1, device=spec.device, dtype=th.float32)
    L.call("maua_audio_onset_env_f32", mel.data_ptr(), env.data_ptr(), scalar.data_ptr(), spec.shape[0], 128,
           1 + N_FFT // (2 * HOP), 1e-10, 80.0, L.stream_ptr(spec.device))
    return env


def resample(x, num):
    """scipy.signal.resample along axis 0."""
    x = _to_dev(x)
    flat = x.reshape(x.shape[0], -1)
    n_in, c = flat.shape
    y = th.empty((num, c), device=x.device, dtype=th.float32)
    ws = th.empty((n_in + num) * c + 2 * (n_in // 2 + num // 2 + 2) * c + 2, device=x.device, dtype=th.float64)
    L.call("maua_resample_f32", flat.data_ptr(), y.data_ptr(), ws.data_ptr(), n_in, num, c, L.stream_ptr(x.device))
    return y.reshape((num,) + tuple(x.shape[1:]))


def _resample_clipped(x, n_frames):
    y = resample(x, n_frames)
    L.call("maua_clip_to_range_f32", x.data_ptr(), x.numel(), y.data_ptr(), y.numel(), L.stream_ptr(x.device))
    return y


# ---------------------------------------------------------------------------------------------------------------
# reference API
# ---------------------------------------------------------------------------------------------------------------

MM_FRAME, MM_HOP = 2048, 441


def onset_strength_mm(y, sr, fmin, fmax):
    """The madmom half of signal.py:52-67: FramedSignal(2048, hop 441) -> STFT(circular_shift) -> |.| ->
    24-bands-per-octave LogarithmicFilterbank(fmin, fmax) -> spectral_diff + spectral_flux + superflux + complex_flux +
    modified_kullback_leibler, one value per hop (csrc/audio.cu, madmom restated — parity unpinned)."""
    y = _to_dev(y)
    n = y.numel()
    T = int(math.ceil(n / float(MM_HOP)))
    n_bins = MM_FRAME // 2
    spec = th.empty((T, n_bins + 1, 2), device=y.device, dtype=th.float32)
    ws = th.empty(T * MM_FRAME, device=y.device, dtype=th.float32)
    L.call("maua_audio_stft_mm_f32", y.data_ptr(), n, spec.data_ptr(), ws.data_ptr(), MM_FRAME, MM_HOP, T,
           L.stream_ptr(y.device))
    fb, lo, hi = filters.log_filterbank(sr, n_bins, 24, fmin, fmax)
    fb_d, lo_d, hi_d = (th.from_numpy(a).to(y.device) for a in (fb, lo, hi))
    # madmom's SpectrogramDifference: frames to look back = the hop count that spans half of the window's half-width
    win = np.hanning(MM_FRAME)
    diff_frames = int(max(1, round((MM_FRAME / 2 - np.argmax(win > 0.5 * win.max())) / MM_HOP)))
    onset = th.empty(T, device=y.device, dtype=th.float32)
    filt = th.empty((T, fb.shape[0]), device=y.device, dtype=th.float32)
    lgd = th.empty((T, n_bins), device=y.device, dtype=th.float32)
    L.call("maua_audio_onsets_mm_f32", spec.data_ptr(), fb_d.data_ptr(), lo_d.data_ptr(), hi_d.data_ptr(),
           onset.data_ptr(), filt.data_ptr(), lgd.data_ptr(), T, n_bins + 1, n_bins, fb.shape[0], diff_frames,
           L.stream_ptr(y.device))
    return onset


def onsets(audio, sr, n_frames, margin=8, fmin=20, fmax=8000, smooth=1, clip=100, power=1, type="mm"):
    """signal.py:31-73: percussive separation -> onset strength (librosa flavour "rosa" or madmom flavour "mm", the
    reference default) -> Fourier resample to n_frames -> causal-0 gaussian -> percentile clip -> power."""
    y_perc = percussive(audio, margin=margin)
    if type == "rosa":
        onset = onset_strength(y_perc, sr, fmin, fmax)
    elif type == "mm":
        onset = onset_strength_mm(y_perc, sr, fmin, fmax)
    else:
        raise ValueError(f"onsets: unknown type {type!r} (expected 'rosa' or 'mm')")
    onset = _resample_clipped(onset, n_frames)
    onset = gaussian_filter(onset, smooth, causal=0)
    onset = percentile_clip(onset, clip)
    if power != 1:
        onset = onset ** power
    return _out(onset)


def rms(y, sr, n_frames, fmin=20, fmax=8000, smooth=180, clip=50, power=6):
    """signal.py:76-99"""
    import scipy.signal as signal

    y = _to_dev(y)
    sos = signal.butter(12, [fmin, fmax], "bp", fs=sr, output="sos")
    sos_d = th.from_numpy(np.ascontiguousarray(sos, dtype=np.float64)).to(y.device)
    y_filt = th.empty_like(y)
    L.call("maua_sosfilt_f32", y.data_ptr(), y_filt.data_ptr(), y.numel(), sos_d.data_ptr(), sos.shape[0],
           L.stream_ptr(y.device))
    spec = stft(y_filt)
    r = th.empty(spec.shape[0], device=y.device, dtype=th.float32)
    L.call("maua_audio_rms_f32", spec.data_ptr(), r.data_ptr(), spec.shape[0], spec.shape[1], N_FFT, L.stream_ptr(y.device))
    r = _resample_clipped(r, n_frames)
    r = gaussian_filter(r, smooth, causal=0.05)
    r = percentile_clip(r, clip)
    return _out(r ** power)


def raw_chroma(audio, sr, type="cens", nearest_neighbor=True):
    """signal.py:102-133 -> [12, n_stft_frames] (STFT chroma front-end for every `type`, see module docstring)."""
    spec = stft(audio)
    raw = _filterbank(spec, filters.chroma(sr, N_FFT))
    T = raw.shape[0]
    if type in ("cens", "deep", "clp"):
        ch = th.empty_like(raw)
        ws = th.empty_like(raw)
        L.call("maua_audio_cens_f32", raw.data_ptr(), ch.data_ptr(), ws.data_ptr(), T, 12, 43, L.stream_ptr(raw.device))
    else:  # "cqt" / "stft": max-normalised chroma
        ch = raw / raw.amax(1, keepdim=True).clamp_min(1e-30)
    if nearest_neighbor and T > 2:
        k = int(min(T - 1, 2 * math.ceil(math.sqrt(T - 2 + 1)), 512))
        rows = min(T, 148 * 2)
        scratch = th.empty(rows * T, device=raw.device, dtype=th.float32)
        out = th.empty_like(ch)
        L.call("maua_audio_nn_filter_f32", ch.data_ptr(), out.data_ptr(), scratch.data_ptr(), rows, T, 12, k,
               L.stream_ptr(raw.device))
        ch = out
    return ch.t()


def chroma(audio, sr, n_frames, margin=16, type="cens", notes=12):
    """signal.py:136-156"""
    y_harm = harmonic(audio, margin=margin)
    ch = raw_chroma(y_harm, sr, type=type).t().contiguous()
    ch = resample(ch, n_frames)
    notes_indices = th.argsort(th.median(ch, dim=0).values)[:notes]
    ch = ch[:, notes_indices]
    return _out(ch / ch.sum(1)[:, None])


def normalize(signal):
    """signal.py:243-254"""
    signal -= signal.min()
    signal /= signal.max()
    return signal


def percentile(signal, p):
    """signal.py:257-270"""
    k = 1 + round(0.01 * float(p) * (signal.numel() - 1))
    return signal.view(-1).kthvalue(k).values.item()


def percentile_clip(signal, p):
    """signal.py:273-292"""
    x = _to_dev(signal)
    y = th.empty_like(x)
    L.call("maua_percentile_clip_f32", x.data_ptr(), y.data_ptr(), x.numel(), float(p), 1.0, L.stream_ptr(x.device))
    return y


def compress(signal, threshold, ratio, invert=False):
    """signal.py:295-311"""
    if invert:
        signal[signal < threshold] *= ratio
    else:
        signal[signal > threshold] *= ratio
    return normalize(signal)


def expand(signal, threshold, ratio, invert=False):
    return compress(signal, threshold, ratio, invert)


def gaussian_filter(x, sigma, causal=None):
    """signal.py:319-368: smooth along the first axis with a (optionally causal) gaussian, circular boundary."""
    xd = _to_dev(x)
    T = xd.shape[0]
    inner = xd.numel() // T
    y = th.empty_like(xd)
    mode, cval = 0, 0.0
    if causal is not None:
        mode, cval = (1, float(causal)) if isinstance(causal, float) else (2, 0.0)
    if int(sigma * 4 * SMF) > T:
        print(f"WARNING: Gaussian filter radius ({int(sigma * 4 * SMF)}) is larger than number of frames ({T}).\n\t "
              f"Filter size has been lowered to ({min(int(sigma * 4 * SMF), 3 * T)}). You might want to consider lowering sigma ({sigma}).")
    L.call("maua_gaussian_filter_f32", xd.data_ptr(), y.data_ptr(), T, inner, float(sigma), float(SMF), mode, cval,
           L.stream_ptr(xd.device))
    return y


LOAD_SR = 22050  # librosa.load's default target rate: every frame constant below (N_FFT, HOP, HPSS widths) assumes it


def load_audio(audio_file, offset=0, duration=-1, cache=True):
    """signal.py:371-405 without librosa: .npy (float array + sibling .sr.txt or 22050 Hz) and PCM / float .wav.

    Like `rosa.load(audio_file, offset=offset, duration=duration)` the result is MONO at 22050 Hz: other sample rates
    are converted with a polyphase resampler (scipy.signal.resample_poly; librosa uses soxr/resampy — same band limit,
    different filter, hence "like").  Integer PCM is scaled by 2**(bits-1) (unsigned 8-bit recentred first)."""
    p = Path(audio_file)
    if p.suffix == ".npy":
        audio = np.load(p).astype(np.float32)
        sr_file = p.with_suffix(".sr.txt")
        sr = int(sr_file.read_text()) if sr_file.exists() else LOAD_SR
    elif p.suffix == ".wav":
        from scipy.io import wavfile

        sr, data = wavfile.read(p)
        if data.dtype.kind == "u":      # 8-bit PCM is unsigned, centred on 128
            bits = data.dtype.itemsize * 8
            data = (data.astype(np.float32) - 2.0 ** (bits - 1)) / 2.0 ** (bits - 1)
        elif data.dtype.kind == "i":
            data = data.astype(np.float32) / 2.0 ** (data.dtype.itemsize * 8 - 1)
        audio = data.astype(np.float32)
    else:
        raise L.MauaError(f"cannot decode {p.suffix} without librosa/ffmpeg: convert to .wav or .npy")
    if audio.ndim > 1:
        audio = audio.mean(1)
    start = int(round(offset * sr))
    total = len(audio) / sr
    if duration == -1 or total < duration:   # signal.py:386-389
        duration = total
        if offset != 0:
            duration -= offset
    audio = audio[start:start + int(round(duration * sr))]
    if sr != LOAD_SR:
        from math import gcd

        from scipy.signal import resample_poly

        g = gcd(int(sr), LOAD_SR)
        audio = resample_poly(audio.astype(np.float64), LOAD_SR // g, int(sr) // g).astype(np.float32)
        sr = LOAD_SR
    return audio, sr, duration
