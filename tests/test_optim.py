"""Optimizers step-for-step against torch.optim."""
import copy

import pytest
import torch

from tiny_deepspeed_b200 import SGD, AdamW


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4))


def _run(ours_cls, ours_kw, ref_cls, ref_kw, steps=6):
    m1, m2 = _model(), _model()
    o1 = ours_cls(m1.named_parameters(), **ours_kw)
    o2 = ref_cls(m2.parameters(), **ref_kw)
    x = torch.randn(5, 8)
    for _ in range(steps):
        for m in (m1, m2):
            m(x).pow(2).sum().backward()
        o1.step()
        o2.step(); o2.zero_grad()
        assert all(p.grad is None for p in m1.parameters())          # step() clears grads (Q4)
    for a, b in zip(m1.parameters(), m2.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    return o1


def test_adam_coupled_matches_torch_adam():
    o = _run(AdamW, dict(lr=1e-2, weight_decay=0.1), torch.optim.Adam, dict(lr=1e-2, weight_decay=0.1))
    assert o.step_count == 6                                          # per-step t, not per tensor (Q3)


def test_adamw_decoupled_matches_torch_adamw():
    _run(AdamW, dict(lr=1e-2, weight_decay=0.1, decoupled=True), torch.optim.AdamW, dict(lr=1e-2, weight_decay=0.1))


def test_amsgrad_really_tracks_max():
    o = _run(AdamW, dict(lr=1e-2, weight_decay=0.0, amsgrad=True), torch.optim.Adam,
             dict(lr=1e-2, weight_decay=0.0, amsgrad=True))
    assert all(st["max_exp_avg_sq"].abs().sum() > 0 for st in o.state.values())


@pytest.mark.parametrize("kw", [dict(lr=0.01), dict(lr=0.01, momentum=0.9), dict(lr=0.01, momentum=0.9, nesterov=True),
                                dict(lr=0.01, momentum=0.9, dampening=0.1, weight_decay=0.01), dict(lr=0.001, maximize=True)])
def test_sgd_matches_torch(kw):
    _run(SGD, kw, torch.optim.SGD, kw)


def test_validation_and_state_dict_roundtrip():
    with pytest.raises(ValueError):
        AdamW(_model().named_parameters(), lr=-1)
    with pytest.raises(ValueError):
        SGD(_model().named_parameters(), lr=0.1, nesterov=True)
    m = _model()
    o = AdamW(m.named_parameters(), lr=1e-2)
    m(torch.randn(3, 8)).sum().backward(); o.step()
    sd = copy.deepcopy(o.state_dict())
    m2 = _model(); o2 = AdamW(m2.named_parameters(), lr=1e-2)
    o2.load_state_dict(sd)
    assert o2.step_count == 1
    for n in o.state:
        torch.testing.assert_close(o.state[n]["exp_avg"], o2.state[n]["exp_avg"])


def test_bf16_params_get_fp32_master():
    m = _model().to(torch.bfloat16)
    o = AdamW(m.named_parameters(), lr=1e-3)
    assert all(st["master"].dtype == torch.float32 for st in o.state.values())
    m(torch.randn(3, 8, dtype=torch.bfloat16)).sum().backward(); o.step()
    for n, p in m.named_parameters():
        torch.testing.assert_close(p.float(), o.state[n]["master"].to(torch.bfloat16).float())
