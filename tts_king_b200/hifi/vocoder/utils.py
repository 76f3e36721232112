"""Helpers the generator shares with the reference's ``hifi/vocoder/utils.py``.

Only the two functions on the inference path are mirrored: ``get_padding`` (reference
``hifi/vocoder/utils.py:36-37``) and ``init_weights`` (``:24-27``).  The plotting / checkpoint-scan
helpers of that file are outside the hot path.
"""


def get_padding(kernel_size, dilation=1):
    """'Same' padding of a dilated, stride-1 Conv1d: ``int((k*d - d) / 2)``."""
    return int((kernel_size * dilation - dilation) / 2)


def init_weights(m, mean=0.0, std=0.01):
    """N(mean, std) on every conv module's ``.weight``.

    As in the reference this runs AFTER ``weight_norm`` and therefore only touches the derived
    ``.weight`` attribute, not ``weight_g``/``weight_v``; it is kept because it
    consumes the global RNG, which is what makes ``torch.manual_seed(s); Generator(h)`` produce the
    same parameters here as there.
    """
    if "Conv" in m.__class__.__name__:
        m.weight.data.normal_(mean, std)
