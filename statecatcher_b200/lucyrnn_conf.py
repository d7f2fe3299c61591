"""LucyRNNConfig — the configuration contract of the hot path.

The drop-in boundary fixes this class completely: ``model.py:232-245`` builds it by keyword and
tests build it positionally, so field names, order, types and defaults have to be those of the
reference dataclass (/root/reference/lucyrnn_conf.py:3-16; tests/test_cpu_abi.py compares the two
field by field when the reference is present).  The class is generated from the table below;
validation of ``kernel_impl`` / ``decay_mode`` stays where the reference has it, in the module
(lucyrnn.py:77-78, 132-133), so that building a config never raises.
"""
from dataclasses import field, make_dataclass

_REQUIRED = object()

# (name, type, default, what the CUDA path does with it)
_CONTRACT = (
    ("input_dim", int, _REQUIRED, "feature width per frame, before frame stacking"),
    ("hidden_dim", int, _REQUIRED, "H: width of every gate, of the carried h and S, and of a layer's output"),
    ("num_layers", int, _REQUIRED, "L: stacked cells; one (h, S) pair is carried per layer"),
    ("vocab_size", int, _REQUIRED, "V: width of output_proj (the CTC / RNN-T emission size)"),
    ("return_last_states", bool, True, "False: forward returns the logits alone"),
    ("kernel_impl", str, "native", "'native' and 'triton' are both accepted and both run the sm_100a kernels"),
    ("is_training", bool, True, "True: segment path (S scan restarts from zero); False: streaming step path"),
    ("fused_ops", bool, False, "True: one [6H, H] gate projection; False: six [H, H] projections plus W_h"),
    ("layer_norm", bool, True, "False swaps the four LayerNorms for identities (the configuration model.py wires)"),
    ("stack_order", int, 1, "consecutive frames concatenated into one input step; the remainder is trimmed"),
    ("decay_mode", str, "learned", "'learned' (sigmoid gate) or 'prefix_sum' (fixed exp(-lambda t) decay)"),
    ("lambda_decay", float, 0.001, "lambda of the 'prefix_sum' mode"),
)

LucyRNNConfig = make_dataclass(
    "LucyRNNConfig",
    [(n, t) if d is _REQUIRED else (n, t, field(default=d)) for n, t, d, _ in _CONTRACT],
    module=__name__)
LucyRNNConfig.__doc__ = "Configuration of LucyRNN.\n\n" + "\n".join(
    f"    {n} ({t.__name__}{'' if d is _REQUIRED else f', default {d!r}'}): {doc}" for n, t, d, doc in _CONTRACT)
