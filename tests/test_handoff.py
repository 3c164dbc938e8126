"""On-device hand-off of the features (SURVEY.md 8f NEXT-2): the reference validates through a temp WAV
(base_tts.py:821-827 -> stt_validator.py:116-148: file -> decode -> resample -> WhisperFeatureExtractor on the host ->
features to the GPU).  Here the [n, 80, 3000] tensor rho_b200_validate leaves in HBM IS the encoder's
`input_features`: same layout, dtype and normalisation, consumed in place.

The STT weights are not in this image, so the encoder is a small randomly initialised WhisperModel (transformers,
seeded): the check is that the encoder sees the same thing either way -- hidden states from the in-place features
against hidden states from the reference chain (oracle post-process -> oracle resample -> transformers'
WhisperFeatureExtractor on the host -> .to(device))."""
import numpy as np
import pytest
import torch


@pytest.mark.gpu
def test_features_feed_a_whisper_encoder_in_place(cuda_device):
    tr = pytest.importorskip("transformers")
    import oracle
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    lens = [240000, 100001, 36000, 480000]
    clips = [c.numpy() for c in synth.make_clips(lens, 41)]
    p = R.make_params()
    rb = R.RaggedBatch.from_list([torch.from_numpy(c) for c in clips], cuda_device)
    v = R.validate_batch(rb, p, n_mels=80, pad_to_30s=True)
    assert v.mel.is_cuda and v.mel.shape == (4, 80, 3000) and v.mel.dtype == torch.float32 and v.mel.is_contiguous()

    cfg = tr.WhisperConfig(d_model=64, encoder_layers=2, decoder_layers=1, encoder_attention_heads=2,
                           decoder_attention_heads=2, encoder_ffn_dim=128, decoder_ffn_dim=128, num_mel_bins=80,
                           vocab_size=1000, max_source_positions=1500, max_target_positions=64, pad_token_id=0,
                           bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    torch.manual_seed(0)
    enc = tr.WhisperModel(cfg).eval().to(cuda_device).encoder
    with torch.no_grad():
        got = enc(input_features=v.mel).last_hidden_state          # no host copy, no re-layout

    # the reference chain on the host
    fe = tr.WhisperFeatureExtractor(feature_size=80)
    c = oracle.derive_constants()
    feats = []
    for x in clips:
        y = oracle.smooth_segment_join([x], c).audio
        w = oracle.resample(y)
        feats.append(fe(w, sampling_rate=16000, return_tensors="np")["input_features"][0])
    ref_in = torch.from_numpy(np.stack(feats)).to(cuda_device)
    with torch.no_grad():
        want = enc(input_features=ref_in).last_hidden_state
    assert float((v.mel - ref_in).abs().max()) <= 1e-4
    err = float((got - want).abs().max())
    scale = float(want.abs().max())
    assert err <= 2e-3 * max(1.0, scale), (err, scale)


class _FakeTokenizer:
    """ids -> words; enough to exercise the plumbing without the (absent) Whisper vocabulary."""
    WORDS = ["hello", "world", "the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog"]

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(self.WORDS[int(t) % len(self.WORDS)] for t in row if int(t) > 2) for row in ids]


@pytest.mark.gpu
def test_validate_audio_text_match_tensor(cuda_device):
    """The tensor-taking sibling of validate_audio_text_match (stt_validator.py:235-259): same return contract, features
    computed on the B200 and consumed in place by a (randomly initialised) transformers Whisper model."""
    tr = pytest.importorskip("transformers")
    import oracle
    import rho_tts_b200 as R
    from rho_tts_b200 import synth
    cfg = tr.WhisperConfig(d_model=64, encoder_layers=2, decoder_layers=1, encoder_attention_heads=2,
                           decoder_attention_heads=2, encoder_ffn_dim=128, decoder_ffn_dim=128, num_mel_bins=80,
                           vocab_size=1000, max_source_positions=1500, max_target_positions=64, pad_token_id=0,
                           bos_token_id=1, eos_token_id=2, decoder_start_token_id=1, suppress_tokens=None,
                           begin_suppress_tokens=None)
    torch.manual_seed(0)
    model = tr.WhisperForConditionalGeneration(cfg).eval().to(cuda_device)
    model.generation_config.suppress_tokens = None
    model.generation_config.begin_suppress_tokens = None
    model.generation_config.forced_decoder_ids = None
    seen = {}
    model.model.encoder.register_forward_pre_hook(
        lambda mod, args, kwargs: seen.__setitem__("x", (kwargs.get("input_features", args[0] if args else None)).detach().clone()),
        with_kwargs=True)
    x = synth.make_clip_block(1, 120000, 17)[0]                      # 5 s at 24 kHz
    gk = dict(max_new_tokens=6, do_sample=False)
    sim_fn = lambda a, b: 1.0 if b else 0.0                          # noqa: E731
    ok, sim, text = R.validate_audio_text_match_tensor(x, 24000, "hello world", 0.85, model=model,
                                                       tokenizer=_FakeTokenizer(), similarity_fn=sim_fn, generate_kwargs=gk)
    assert isinstance(ok, bool) and isinstance(sim, float) and (text is None or isinstance(text, str))
    assert text is not None and (ok, sim) == (True, 1.0)
    # what the encoder saw is what the reference chain computes on the host for the same samples
    fe = tr.WhisperFeatureExtractor(feature_size=80)
    want = fe(oracle.resample(x.numpy()), sampling_rate=16000, return_tensors="np")["input_features"]
    assert seen["x"].is_cuda and tuple(seen["x"].shape) == (1, 80, 3000)
    assert float((seen["x"].float().cpu() - torch.from_numpy(want)).abs().max()) <= 1e-4
    # other input rates go through the any-ratio resampler; 16 kHz is taken as is
    f16 = R.whisper_features(torch.from_numpy(oracle.resample(x.numpy())), 16000)
    assert float((f16.cpu() - torch.from_numpy(want)).abs().max()) <= 1e-4
    f22 = R.whisper_features(x[:44100], 22050)
    assert tuple(f22.shape) == (1, 80, 3000) and torch.isfinite(f22).all()
    # a failing transcription is "validation skipped", not "invalid" (stt_validator.py:251-253)

    class Broken:
        config = cfg
        dtype = torch.float32

        def generate(self, **kw):
            raise RuntimeError("decoder exploded")
    assert R.validate_audio_text_match_tensor(x, 24000, "hello", model=Broken(), tokenizer=_FakeTokenizer(),
                                              similarity_fn=sim_fn) == (True, 0.0, None)
    with pytest.raises(RuntimeError):
        R.validate_audio_text_match_tensor(x, 24000, "hello")
