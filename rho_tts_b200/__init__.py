"""rho_tts_b200 -- B200-native (sm_100a) audio post-processing + validation front end for rho-tts.

Drop-in for the data-parallel hot path of rhofield/rho-tts: silence trim,
DC removal, fades, crossfade joins, sound-decay check, 24k->16k resampling, Whisper log-mel and
speaker-embedding cosine, behind the reference's own BaseTTS method signatures (`B200AudioMixin`)
plus batched entry points over a ragged HBM layout.  The arithmetic lives in librho_b200.so
(C ABI: include/rho_b200.h); this package is the host mirror and has no CPU fallback.
"""
from .ragged import RaggedBatch, ALIGN                                     # noqa: F401
from .batch import (make_params, params_from_tts, trim_scan_batch, join_batch, post_process_batch,  # noqa: F401
                    resample_batch, resample_any_batch, logmel_batch, mel_project, stft_power_tc, qwen_post_process_batch, qwen_pipeline_batch, qwen_validate_batch, pitch_shift_batch, mfcc_stats_batch, pcm16_batch, write_wav, cosine_batch, validate_batch, validate_host, validate_host_ragged, host_item_layout, HostRaggedOutput,
                    ValidatePlan, JoinOutput, ValidateOutput, REC_DTYPE, SEG_DTYPE)
from .mixin import B200AudioMixin, B200QwenAudioMixin, make_b200_provider, register_b200_providers   # noqa: F401
from . import _lib                                                         # noqa: F401
from .validation import whisper_features, transcribe_tensor, validate_audio_text_match_tensor   # noqa: F401
from . import speaker                                                      # noqa: F401  (resemblyzer's front end, NEXT-3)
from . import dist                                                         # noqa: F401  (sharding, record gather, NUMA binding)

__version__ = "0.1.0"
