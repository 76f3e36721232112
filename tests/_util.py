"""Shared helpers for the test-suite."""
import os
import warnings

import numpy as np
import torch

from oracle import fixtures as fx
from oracle.common import config_from_h

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def stored_state(g, prefix="sd."):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def make_generator(cfg, seed=1234, fold=True, precision="fp32"):
    """The package's Generator(h) under the reference's seed (hifi.seed, config.yaml:23)."""
    from tts_king_b200.hifi.models import Generator

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(seed)
        m = Generator(fx.make_h(cfg), precision=precision)
        if fold:
            import contextlib
            import io

            with contextlib.redirect_stdout(io.StringIO()):
                m.remove_weight_norm()
    return m.eval()


def np_state(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}
