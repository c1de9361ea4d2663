"""Pins the oracle's explicit-uniform sampler to torch.multinomial (== Categorical.sample, the call the
reference makes at long_term_attention_gibbs.py:204-205 / long_term_attention.py:230-232)."""
import pytest
import torch

from oracle import ltm_oracle as O


@pytest.mark.parametrize("ncat,power,zeros", [(127, 1, False), (128, 1, False), (127, 8, False), (5, 1, False),
                                               (127, 1, True), (128, 3, True)])
def test_inverse_cdf_matches_multinomial(ncat, power, zeros):
    torch.manual_seed(ncat * 7 + power)
    p = torch.rand(3, ncat) ** power
    if zeros:
        p[:, ::3] = 0
    S = 50000
    torch.manual_seed(99)
    ref = torch.multinomial(p, S, replacement=True)
    torch.manual_seed(99)
    u = torch.rand(3, S, dtype=torch.float64)
    assert torch.equal(O.inverse_cdf_sample(p, u), ref)


def test_rng_accounting_of_a_sticky_call():
    """A sticky reference call with B=1 consumes 512 used + 512 discarded fp64 draws."""
    torch.manual_seed(5)
    p = torch.rand(1, 127)
    d = torch.distributions.Categorical(p)
    d.sample((512,))
    torch.distributions.Categorical(torch.ones(1)).sample((512, 1))
    nxt = torch.rand(1, dtype=torch.float64)
    torch.manual_seed(5)
    torch.rand(1, 127)
    allu = torch.rand(1025, dtype=torch.float64)
    assert nxt.item() == allu[1024].item()
