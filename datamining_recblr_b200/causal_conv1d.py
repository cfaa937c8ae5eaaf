"""Drop-in for the `causal_conv1d` dependency as the reference calls it (RecBLR.py:188-193):
causal_conv1d_fn(x=[B, C, T] with channel-last strides, weight=[C, W], bias=[C], activation="silu")."""
from .ops import causal_conv1d_channel_last

__all__ = ["causal_conv1d_fn"]


def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                     final_states_out=None, activation=None):
    """x: (batch, dim, seqlen); weight: (dim, width); bias: (dim,); activation in (None, "silu", "swish").
    Returns (batch, dim, seqlen) with the same (channel-last) strides, like the upstream wheel."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    if seq_idx is not None or initial_states is not None or return_final_states:
        raise NotImplementedError("seq_idx / initial_states / return_final_states are not used by RecBLR")
    y = causal_conv1d_channel_last(x.transpose(1, 2), weight, bias, silu=activation is not None)
    return y.transpose(1, 2)
