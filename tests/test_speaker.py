"""Speaker-encoder front end (SURVEY.md 8f NEXT-3): resemblyzer's normalize_volume / compute_partial_slices /
wav_to_mel_spectrogram / embed_utterance pooling, reached from base_tts.py:326-347.  resemblyzer and librosa are not in
this image: the oracle (oracle/speaker.py, parity unpinned against resemblyzer itself) is pinned on transformers' port of
librosa's spectrogram (tests/golden/golden_speaker_v1.npz) and the CUDA path is compared with the oracle.

Tolerance of the 40-band POWER spectrogram (no log, values span ten decades): |got - want| <= 1e-4 |want| + 2e-6 max|want|
per clip -- 1e-4 relative as north_star asks for floating point, plus the fp32 FFT's noise floor relative to the clip's
strongest band (the numpy oracle itself is 4e-7 of the maximum away from the float64 golden vectors)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from speaker_inputs import SPEAKER_LENGTHS, speaker_input  # noqa: E402

from oracle import speaker as OS  # noqa: E402


def _close(got, want, rtol=1e-4, floor=2e-6):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.size == 0:
        return 0.0
    tol = rtol * np.abs(want) + floor * np.abs(want).max()
    err = np.abs(got - want)
    assert (err <= tol).all(), float((err / np.maximum(tol, 1e-300)).max())
    return float(err.max())


# ------------------------------------------------------------------------------------------------ CPU: oracle, host logic
def test_oracle_mel_matches_golden():
    g = np.load(os.path.join(HERE, "golden", "golden_speaker_v1.npz"))
    assert np.array_equal(OS.mel_filterbank(40, 400, 16000), g["filterbank"])
    for i, n in enumerate(SPEAKER_LENGTHS):
        if n == 0:
            assert OS.wav_to_mel_spectrogram(np.zeros(0, np.float32)).shape == (1, 40)
            continue
        m = OS.wav_to_mel_spectrogram(speaker_input(i))
        assert m.shape == (1 + n // 160, 40) and m.dtype == np.float32
        _close(m[::3] if m.shape[0] > 64 else m, g[f"mel_sub{i}"])


def test_partial_slices_host_function_matches_oracle():
    from rho_tts_b200 import speaker as SP
    rng = np.random.default_rng(5)
    ns = list(range(0, 400)) + [25599, 25600, 25601, 31519, 31520, 31521, 37920, 160000, 480000, 2 ** 31 - 2]
    ns += [int(v) for v in rng.integers(0, 2_000_000, 300)]
    for rate, cov in ((1.3, 0.75), (1.0, 0.5), (2.5, 1.0), (0.7, 0.9)):
        for n in ns:
            ws, ms = OS.compute_partial_slices(n, rate, cov)
            gw, gm = SP.compute_partial_slices(n, rate, cov)
            assert gm == ms and gw == ws, (n, rate, cov)
            assert SP.partial_count(n, rate, cov) == (len(ms), ws[-1].stop)
        cnt, end = SP.partial_counts(np.asarray(ns), rate, cov)
        assert [(int(a), int(b)) for a, b in zip(cnt, end)] == [SP.partial_count(n, rate, cov) for n in ns]
    assert SP.frame_step_of(1.3) == 77 == OS.frame_step_of(1.3)
    with pytest.raises(AssertionError):
        SP.frame_step_of(0.5)                       # "The rate is too low"
    lib = SP._lib.load()
    assert lib.rho_b200_spk_slices(-1, 77, 0.75, None) < 0 and lib.rho_b200_spk_slices(10, 161, 0.75, None) < 0
    assert lib.rho_b200_spk_slices(10, 77, 0.0, None) < 0


def test_speaker_filterbank_table():
    from rho_tts_b200 import _lib
    lib = _lib.load()
    a = np.zeros(40 * 201, np.float32)
    assert lib.rho_b200_host_table(2, 40, a.ctypes.data, a.size) == a.size
    assert np.array_equal(a.reshape(40, 201), OS.mel_filterbank(40, 400, 16000))


def test_oracle_normalize_volume_contract():
    x = speaker_input(1)                                # quiet: raised to -30 dBFS
    y = OS.normalize_volume(x, -30, increase_only=True)
    rms = np.sqrt(np.mean(y.astype(np.float64) ** 2))
    assert abs(20 * np.log10(rms) + 30) < 1e-3 and y.dtype == np.float32
    loud = speaker_input(0)
    assert OS.normalize_volume(loud, -30, increase_only=True) is not None
    assert OS.volume_gain(loud, -30, increase_only=True) == 1.0
    assert OS.volume_gain(loud, -30, decrease_only=True) < 1.0
    with pytest.raises(ValueError):
        OS.normalize_volume(x, -30, True, True)


# ------------------------------------------------------------------------------------------------ GPU
def _batch(device, idx=None):
    import rho_tts_b200 as R
    idx = list(range(len(SPEAKER_LENGTHS))) if idx is None else idx
    clips = [speaker_input(i) for i in idx]
    return clips, R.RaggedBatch.from_list([torch.from_numpy(c) for c in clips], device)


@pytest.mark.gpu
def test_normalize_volume_vs_oracle(cuda_device):
    from rho_tts_b200 import speaker as SP
    clips, rb = _batch(cuda_device)
    for kw in (dict(increase_only=True), dict(decrease_only=True), dict()):
        g = SP.volume_gains(rb, -30.0, **kw).cpu().numpy()
        out = SP.normalize_volume(rb, -30.0, **kw)
        for i, c in enumerate(clips):
            want_g = OS.volume_gain(c, -30, **kw)
            assert abs(g[i] - want_g) <= 2e-6 * want_g, (i, kw, g[i], want_g)
            want = OS.normalize_volume(c, -30, **kw)
            got = out.clip(i).cpu().numpy()
            assert got.shape == want.shape
            if want.size:
                assert float(np.abs(got - want).max()) <= 3e-6 * float(np.abs(want).max())
            if want_g == 1.0:
                assert np.array_equal(got, c)           # untouched clips are bit-identical
    with pytest.raises(ValueError):
        SP.volume_gains(rb, -30.0, True, True)
    # in place, and an all-zero clip: the reference's gain is +inf (0 * inf = NaN); an empty one is left alone
    import rho_tts_b200 as R
    z = R.RaggedBatch.from_list([torch.zeros(1000), torch.zeros(0), torch.from_numpy(clips[1])], cuda_device)
    g = SP.volume_gains(z, -30.0, increase_only=True, out=z).cpu().numpy()
    assert np.isinf(g[0]) and g[1] == 1.0
    assert torch.isnan(z.clip(0)).all()
    assert float(np.abs(z.clip(2).cpu().numpy() - OS.normalize_volume(clips[1], -30, increase_only=True)).max()) <= 1e-7


@pytest.mark.gpu
def test_mel_spectrogram_vs_oracle_and_golden(cuda_device):
    from rho_tts_b200 import speaker as SP
    clips, rb = _batch(cuda_device)
    mel, foff = SP.wav_to_mel_spectrogram(rb)
    assert mel.shape == (int(foff[-1]), 40) and mel.dtype == torch.float32
    g = np.load(os.path.join(HERE, "golden", "golden_speaker_v1.npz"))
    m_all = mel.cpu().numpy()
    for i, c in enumerate(clips):
        got = m_all[foff[i]:foff[i + 1]]
        want = OS.wav_to_mel_spectrogram(c)
        _close(got, want)
        if len(c):
            _close(got[::3] if got.shape[0] > 64 else got, g[f"mel_sub{i}"])
    # with a gain per clip: the spectrogram of the normalised waveform
    gain = SP.volume_gains(rb, -30.0, increase_only=True)
    mel2, foff2 = SP.wav_to_mel_spectrogram(rb, gain=gain)
    m2 = mel2.cpu().numpy()
    for i, c in enumerate(clips):
        _close(m2[foff2[i]:foff2[i + 1]], OS.wav_to_mel_spectrogram(OS.normalize_volume(c, -30, increase_only=True)), rtol=1e-4, floor=4e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("rate,cov", [(1.3, 0.75), (2.0, 0.5)])
def test_partial_mels_vs_oracle(cuda_device, rate, cov):
    from rho_tts_b200 import speaker as SP
    clips, rb = _batch(cuda_device)
    partials, poff = SP.partial_mels(rb, rate, cov)
    assert partials.shape == (int(poff[-1]), 160, 40)
    p_all = partials.cpu().numpy()
    assert np.isfinite(p_all).all()
    for i, c in enumerate(clips):
        want = OS.partial_mels(c, rate, cov)
        got = p_all[poff[i]:poff[i + 1]]
        assert got.shape == want.shape, (i, got.shape, want.shape)
        scale = max(float(want.max()), 1e-30)
        for j in range(want.shape[0]):
            tol = 1e-4 * np.abs(want[j].astype(np.float64)) + 2e-6 * scale
            assert (np.abs(got[j].astype(np.float64) - want[j]) <= tol).all(), (i, j)


class _Lstm(torch.nn.Module):
    """resemblyzer's architecture (voice_encoder.py: LSTM 40 -> 256 x 3, Linear 256 -> 256, ReLU, L2 normalisation),
    randomly initialised: the pretrained weights are not in this image."""

    def __init__(self):
        super().__init__()
        self.lstm = torch.nn.LSTM(40, 256, 3, batch_first=True)
        self.linear = torch.nn.Linear(256, 256)
        self.relu = torch.nn.ReLU()

    def forward(self, mels):
        _, (hidden, _) = self.lstm(mels)
        e = self.relu(self.linear(hidden[-1]))
        return e / torch.norm(e, dim=1, keepdim=True)


@pytest.mark.gpu
def test_pool_and_similarity_vs_oracle(cuda_device):
    import rho_tts_b200 as R
    from rho_tts_b200 import speaker as SP
    rng = np.random.default_rng(3)
    counts = [1, 4, 12, 2, 7]
    pe = rng.normal(0, 1, (sum(counts), 256)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    got = SP.pool_partials(torch.from_numpy(pe).to(cuda_device), off).cpu().numpy()
    for i in range(len(counts)):
        want = OS.pool_partials(pe[off[i]:off[i + 1]])
        assert float(np.abs(got[i] - want).max()) <= 2e-7
    # the whole of _compute_speaker_similarity on 24 kHz clips, the encoder being the same random LSTM on both sides
    import oracle
    from rho_tts_b200 import synth
    torch.manual_seed(1)
    enc = _Lstm().eval()
    lens = [72000, 48001, 240000, 30000]
    clips = [c.numpy() for c in synth.make_clips(lens, 77)]
    clips[1] = (clips[1] * 0.01).astype(np.float32)
    ref_emb = rng.normal(0, 1, 256).astype(np.float32)
    want = []
    for c in clips:
        w16 = oracle.resample(c)
        w16 = OS.normalize_volume(w16, -30, increase_only=True)
        e = OS.embed_utterance(w16, lambda m: enc(torch.from_numpy(m)).detach().numpy())
        want.append(OS.speaker_similarity(ref_emb, e))
    enc_d = _Lstm().eval()
    enc_d.load_state_dict(enc.state_dict())
    enc_d = enc_d.to(cuda_device)
    rb = R.RaggedBatch.from_list([torch.from_numpy(c) for c in clips], cuda_device)
    got = SP.speaker_similarity(rb, torch.from_numpy(ref_emb), enc_d, sample_rate=24000).cpu().numpy()
    assert got.shape == (4,)
    assert float(np.abs(got - np.asarray(want, dtype=np.float32)).max()) <= 1e-4
    with pytest.raises(RuntimeError):
        SP.speaker_similarity(rb, torch.from_numpy(ref_emb), enc_d, sample_rate=22050)

    # the same through the method the reference calls (base_tts.py:326-347), opted in on the mixin
    class Fake(R.B200AudioMixin):
        sample_rate = 24000
        b200_speaker_front_end = True
        voice_encoder = enc_d
        reference_embedding = ref_emb
    enc_d.device = cuda_device                       # resemblyzer's VoiceEncoder carries its device like this
    tts = Fake()
    for i, c in enumerate(clips[:2]):
        s = tts._compute_speaker_similarity(torch.from_numpy(c).unsqueeze(0))
        assert isinstance(s, np.float32) and abs(float(s) - float(want[i])) <= 1e-4


@pytest.mark.gpu
def test_speaker_entry_points_reject_bad_arguments(cuda_device):
    from rho_tts_b200 import _lib
    from rho_tts_b200 import speaker as SP
    h = _lib.Handle.get(0)
    clips, rb = _batch(cuda_device, [0, 1])
    P = lambda t: None if t is None else t.data_ptr()            # noqa: E731
    assert h.lib.rho_b200_spk_mel(h.ptr, P(rb.data), P(rb.offsets), P(rb.lengths), 4, 2, rb.max_len, 77, 0.75, 0, None,
                                  None, None, None, None, None) == -1            # no output asked for
    mel = torch.empty((400, 40), device=cuda_device)
    assert h.lib.rho_b200_spk_mel(h.ptr, P(rb.data), P(rb.offsets), P(rb.lengths), 4, 2, rb.max_len, 0, 0.75, 0, None,
                                  P(mel), None, None, None, None) == -1          # frame_step, and mel without offsets
    g = torch.empty(2, device=cuda_device)
    assert h.lib.rho_b200_normalize_volume(h.ptr, P(rb.data), P(rb.offsets), P(rb.lengths), 4, 2, rb.max_len, -30.0, 3, None,
                                           None, P(g), None, 0, None) == -1      # mode
    assert h.lib.rho_b200_normalize_volume(h.ptr, P(rb.data), P(rb.offsets), P(rb.lengths), 4, 2, rb.max_len, -30.0, 1, None,
                                           None, P(g), None, 0, None) == -4      # RHO_ERR_WORKSPACE
    assert SP.partial_mels(SP.RaggedBatch.from_list([], cuda_device))[0].shape == (0, 160, 40)
