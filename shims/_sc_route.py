"""Route the criterion train.py builds — ``nn.CTCLoss(blank=blank_id, zero_infinity=True)``,
train.py:142, called at model.py:71 — to the sm_100a CTC kernels WITHOUT touching train.py.

Every shim module imports this file, so by the time train.py reaches ``setup_criterion`` (it imports
``model`` -> ``lucyrnn`` / ``lucyrnn_conf`` / ``lucyrnn_triton`` at its top) ``torch.nn.CTCLoss`` is a subclass
of torch's own class whose forward sends CUDA inputs through ``statecatcher_b200.ctc_loss`` (same
arguments, Python-list lengths included, (T,B,V) log-probs as model.py:70 passes them — the kernels
fold the idempotent log-softmax in) and everything else (CPU tensors, dtypes the kernels do not take)
through torch's implementation.  ``isinstance(x, nn.CTCLoss)`` and pickling by attribute keep working.

Switch: ``SC_SHIM_CTC=0`` in the environment leaves torch's class alone (default: on).
"""
import os

import torch
import torch.nn as nn


def install():
    if os.environ.get("SC_SHIM_CTC", "1") == "0":
        return False
    if getattr(nn.CTCLoss, "_statecatcher_b200_routed", False):
        return True
    from statecatcher_b200.ctc import ctc_loss

    torch_ctc = nn.CTCLoss

    class CTCLoss(torch_ctc):
        _statecatcher_b200_routed = True
        _torch_class = torch_ctc

        def forward(self, log_probs, targets, input_lengths, target_lengths):
            if isinstance(log_probs, torch.Tensor) and log_probs.is_cuda and log_probs.dtype in (
                    torch.float32, torch.bfloat16, torch.float16):
                return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank,
                                self.reduction, self.zero_infinity)
            return super().forward(log_probs, targets, input_lengths, target_lengths)

    CTCLoss.__name__ = torch_ctc.__name__
    CTCLoss.__qualname__ = torch_ctc.__qualname__
    CTCLoss.__doc__ = torch_ctc.__doc__
    nn.CTCLoss = CTCLoss
    nn.modules.loss.CTCLoss = CTCLoss
    return True


def uninstall():
    cur = nn.CTCLoss
    if getattr(cur, "_statecatcher_b200_routed", False):
        nn.CTCLoss = cur._torch_class
        nn.modules.loss.CTCLoss = cur._torch_class


install()
