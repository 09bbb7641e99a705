"""sopro_b200 — B200-native engine for the Sopro TTS hot path behind the reference's API.

    from sopro_b200 import SoproTTS          # drop-in for `from sopro import SoproTTS`

Importing the package does not need a GPU; constructing a model does (no CPU fallback)."""
from .config import SoproTTSConfig  # noqa: F401

__version__ = "0.1.0"
__all__ = ["SoproTTS", "SoproTTSConfig"]


def __getattr__(name):  # lazy: keep `import sopro_b200` cheap and GPU-free
    if name == "SoproTTS":
        from .model import SoproTTS

        return SoproTTS
    if name == "PreparedReference":
        from .prefill import PreparedReference

        return PreparedReference
    raise AttributeError(name)
