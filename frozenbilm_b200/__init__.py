"""frozenbilm_b200: B200-native (sm_100a) implementation of FrozenBiLM's masked-LM hot path."""


def install_as_model():
    """Registers frozenbilm_b200.model as the top-level package ``model`` so that the reference's
    ``from model import build_model, get_tokenizer`` (main.py:17, mc.py:19) resolves to this build."""
    import sys

    from . import model as _m

    sys.modules["model"] = _m
    return _m
