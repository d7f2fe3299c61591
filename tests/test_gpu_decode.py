"""GPU: greedy CTC decoder vs the oracle restatement of decoder.py:3-30 — exact (integer work)."""
import numpy as np
import pytest
import torch

from oracle import decoder_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,T,V", [(1, 1, 2), (3, 17, 5), (4, 300, 33), (2, 777, 1024)])
def test_greedy_decode_exact(cuda_device, dtype, B, T, V):
    import statecatcher_b200 as sb
    g = torch.Generator().manual_seed(B * T + V)
    # coarse integer-valued scores: many exact ties and many repeated argmaxes
    x = torch.randint(0, 4, (B, T, V), generator=g).to(dtype)
    lens = [T] + [int(torch.randint(0, T + 1, (1,), generator=g)) for _ in range(B - 1)]
    want = decoder_oracle.ctc_greedy_decode(x.float().numpy(), lens, blank=0)
    got = sb.ctc_greedy_decoder(x.cuda(), torch.tensor(lens).cuda(), blank=0)
    assert got == want
    got2 = sb.ctc_greedy_decoder(x.cuda(), lens, blank=0)          # list lengths (train.py passes a tensor)
    assert got2 == want


def test_greedy_decode_blank_and_repeat_rules(cuda_device):
    import statecatcher_b200 as sb
    seq = [0, 3, 3, 0, 3, 5, 5, 5, 0, 0, 2, 2, 3]          # -> 3 3 5 2 3
    V = 6
    x = torch.full((1, len(seq), V), -5.0)
    for t, k in enumerate(seq):
        x[0, t, k] = 1.0
    assert sb.ctc_greedy_decoder(x.cuda(), [len(seq)], blank=0) == [[3, 3, 5, 2, 3]]
    assert sb.ctc_greedy_decoder(x.cuda(), [5], blank=0) == [[3, 3]]
    assert sb.ctc_greedy_decoder(x.cuda(), [0], blank=0) == [[]]
    assert sb.ctc_greedy_decoder(x.cuda(), [len(seq)], blank=3) == decoder_oracle.ctc_greedy_decode(x.numpy(), [len(seq)], blank=3) == [[0, 0, 5, 0, 2]]
