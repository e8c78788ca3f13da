"""navc-b200: sm_100a kernels behind the API of yangbang18/Non-Autoregressive-Video-Captioning.

    from navc_b200.models import get_model            # reference: models.get_model
    from navc_b200.models.Translator import Translator
    from navc_b200.decoding import generate
"""
from . import _lib  # noqa: F401
from .config import Constants  # noqa: F401
from .models import get_model  # noqa: F401
from .models.Translator import Translator  # noqa: F401
from .decoding import generate  # noqa: F401

__version__ = "0.1.0"
