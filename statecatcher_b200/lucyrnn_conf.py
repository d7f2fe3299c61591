"""LucyRNNConfig — the configuration contract of the hot path.

Same twelve fields, names, order and defaults as the reference dataclass
(/root/reference/lucyrnn_conf.py:3-16) so that ``model.py:232-245`` can build it unchanged.
``kernel_impl`` keeps its two legal values for compatibility; both select the sm_100a
kernels here (there is no Triton path and no PyTorch "native" loop).
"""
from dataclasses import dataclass


@dataclass
class LucyRNNConfig:
    input_dim: int
    hidden_dim: int
    num_layers: int
    vocab_size: int
    return_last_states: bool = True
    kernel_impl: str = "native"   # 'native' | 'triton' accepted; both run the CUDA path
    is_training: bool = True      # True: segment-parallel path; False: streaming step path
    fused_ops: bool = False       # one [6H,H] gate projection instead of six [H,H]
    layer_norm: bool = True
    stack_order: int = 1          # frames stacked per input step
    decay_mode: str = "learned"   # 'learned' | 'prefix_sum'
    lambda_decay: float = 0.001   # only for 'prefix_sum'
