"""CPU oracle for the HiFi-GAN generator path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.  Nothing under ``tts_king_b200/`` does.

Two independent restatements of ``hifi/models.py::Generator`` (reference tree):

* :mod:`oracle.c_oracle`      plain C (``hifigan_oracle.c``), fp32 and fp64 builds
* :mod:`oracle.torch_oracle`  ``torch.nn.functional`` calls on CPU tensors — the
  same ATen/oneDNN dispatch the reference's eager modules reach

Parity pin: the reference ships no golden vectors for this path, so both are
pinned against outputs of the reference itself, generated in the build container
by ``tools/make_golden.py`` and committed under ``tests/golden/``.
"""
