"""Optimizer-step oracle — TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

Lion: train.py:125-131 constructs ``lion_pytorch.Lion(params, lr=, weight_decay=)``.  The package
is a third-party dependency that is NOT under /root/reference, not in requirements.txt and not
installed here (no version to pin) -> **parity unpinned**.  This restates the published rule
(Chen et al. 2023, "Symbolic Discovery of Optimization Algorithms", Algorithm 2), which is also
what the package's ``update_fn`` does, in numpy fp64:

    p <- p * (1 - lr*wd)
    p <- p - lr * sign(beta1*m + (1-beta1)*g)
    m <- beta2*m + (1-beta2)*g

``clip`` restates torch.nn.utils.clip_grad_norm_ (train.py:553): one coefficient
min(1, max_norm / (||g||_2 + 1e-6)) over ALL tensors.
"""
import numpy as np


def clip_coef(grads, max_norm):
    total = np.sqrt(sum(float((np.asarray(g, np.float64) ** 2).sum()) for g in grads))
    return total, min(1.0, max_norm / (total + 1e-6))


def lion_step(p, g, m, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.0):
    """One update of one tensor.  Returns (p_new, m_new, u) with u the pre-sign interpolation, so a
    fp32 checker can leave out the elements whose sign is decided below fp32 resolution."""
    p = np.asarray(p, np.float64); g = np.asarray(g, np.float64); m = np.asarray(m, np.float64)
    b1, b2 = betas
    u = b1 * m + (1.0 - b1) * g
    p_new = p * (1.0 - lr * weight_decay) - lr * np.sign(u)
    m_new = b2 * m + (1.0 - b2) * g
    return p_new, m_new, u
