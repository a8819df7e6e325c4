"""funsor_b200: B200-native (sm_100a) Markov-chain sum-product path for funsor.

The package holds only what the path needs:
  csrc/            hand-written CUDA kernels + the C ABI (include/funsor_b200.h)
  _lib.py          ctypes binding of that ABI (no torch types cross it)
  ops.py           host-side mirror of the reference functions on raw torch CUDA tensors
  gaussian.py      Gaussian-factor (Kalman) mirror
  hmm.py           DiscreteHMM / GaussianHMM / GaussianMRF with the reference's constructor arguments (pyro/hmm.py)
  distributed.py   batch- and time-sharding across the GPUs of one box
  funsor_plugin.py eager patterns registered into funsor's torch backend (when funsor is importable)
There is no CPU fallback: without the shared library or without an sm_100 device every compute
entry point raises.
"""
from . import _lib  # noqa: F401
from .ops import (  # noqa: F401
    async_status,
    chain_adjoint,
    hmm_chunk_summary,
    hmm_chunk_summary_tree,
    hmm_forward,
    hmm_marginals,
    hmm_posterior_sample,
    hmm_viterbi,
    launch_count,
    mixed_sequential_sum_product,
    naive_sequential_sum_product,
    semiring_matmul,
    sequential_sum_product,
)

__version__ = "0.1.0"
