"""tts_king_b200 — B200-native (sm_100a) HiFi-GAN vocoder inference, drop-in for the
``hifi/models.py::Generator`` / ``hifiapi.py::HIFIapi`` path of diff7/tts-king.

    from tts_king_b200.hifi.models import Generator      # was: from hifi.models import Generator
    from tts_king_b200.hifiapi import HIFIapi            # was: from hifiapi import HIFIapi

All arithmetic runs in hand-written CUDA behind the C ABI of ``include/hifigan_b200.h``
(``tts_king_b200/lib/libhifigan_b200.so``); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
