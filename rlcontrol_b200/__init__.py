"""rlcontrol_b200 -- B200-native (sm_100a) sampled-action critic path of RLControl.

Only what the hot path needs: ``csrc/`` (CUDA kernels + the C-ABI of ``include/rlc.h``, built into
``librlc.so``), ``engine`` (torch-tensor wrapper of the C-ABI), ``networks`` (drop-in mirrors of the
reference's ``agents/network`` critic entry points), ``kl_networks`` (drop-in ForwardKLNetwork /
ReverseKLNetwork), ``replaybuffer`` and ``quadrature``.
There is no CPU fallback: importing works anywhere, using it needs the built library and a GPU."""
from . import _lib  # noqa: F401
from ._lib import (ACT_PER_STATE, ACT_SHARED, ADAM_TF, ADAM_TORCH, LAYOUT_IN_OUT, LAYOUT_OUT_IN,  # noqa: F401
                   PREC_AUTO, PREC_BF16, PREC_FP16, PREC_FP16X3, PREC_FP32, TIN, TMID, RlcError)

__all__ = ["Engine", "Critic", "CriticOptimizer", "Mlp"]


def __getattr__(name):
    if name in ("Engine", "Critic", "CriticOptimizer", "Mlp"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
