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
