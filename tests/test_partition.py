"""Golden tables of the reference's partitioner."""
from collections import OrderedDict

import pytest
import torch

from tiny_deepspeed_b200 import partition_tensors
from tiny_deepspeed_b200.parallel import partition_report
from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config


def _toy(n=8):
    return OrderedDict((f"t{i}", torch.empty(4, device="meta")) for i in range(n))


def test_toy_e0_e1():
    t, _ = partition_tensors(_toy(), num_parts=4, evenness_priority=0)
    assert list(t.values()) == [0, 0, 1, 1, 2, 2, 3, 3]
    t, _ = partition_tensors(_toy(), num_parts=4, evenness_priority=1)
    assert list(t.values()) == [0, 1, 2, 3, 3, 3, 3, 3]


def _first_owned(table, ws):
    first = {}
    for name, r in table.items():
        first.setdefault(r, name.replace("transformer.", "").replace(".weight", ""))
    return first


GOLDEN = {
    ("small", 2): {0: "wte", 1: "h.5.mlp.c_proj"},
    ("small", 4): {0: "wte", 1: "h.0.attn.c_attn", 2: "h.5.mlp.c_proj", 3: "h.11.mlp.c_fc"},
    ("small", 8): {0: "wte", 1: "wpe", 2: "h.2.mlp.c_proj", 3: "h.5.mlp.c_fc", 4: "h.8.attn.c_attn",
                   5: "h.10.mlp.c_proj", 6: "lm_head"},
    ("medium", 4): {0: "wte", 1: "h.3.mlp.c_proj", 2: "h.11.mlp.c_proj", 3: "h.19.mlp.c_proj"},
    ("large", 8): {0: "wte", 1: "h.1.mlp.c_proj", 2: "h.6.mlp.c_proj", 3: "h.11.mlp.c_proj", 4: "h.16.mlp.c_proj",
                   5: "h.21.mlp.c_proj", 6: "h.26.mlp.c_proj", 7: "h.31.mlp.c_proj"},
    ("xl", 4): {0: "wte", 1: "h.10.mlp.c_fc", 2: "h.23.mlp.c_fc", 3: "h.36.mlp.c_fc"},
}


@pytest.mark.parametrize("key", list(GOLDEN))
def test_gpt2_golden_boundaries(key):
    name, ws = key
    with torch.device("meta"):
        m = GPT2Model(gpt2_config(name))
    table, _ = partition_tensors(OrderedDict(m.named_parameters()), ranks_map=[f"cuda:{i}" for i in range(ws)],
                                 evenness_priority=0)
    assert _first_owned(table, ws) == GOLDEN[key]


def test_small_ws8_rank7_empty_and_better_strategies():
    with torch.device("meta"):
        m = GPT2Model(gpt2_config("small"))
    params = OrderedDict(m.named_parameters())
    table, _ = partition_tensors(params, num_parts=8)
    rep = partition_report(params, table, 8)
    assert rep["load"][7] == 0 and abs(rep["imbalance"] - 1.90) < 0.02
    for strat, bound in (("contiguous", 1.90), ("balanced", 1.90)):
        t2, _ = partition_tensors(params, num_parts=8, strategy=strat)
        r2 = partition_report(params, t2, 8)
        assert min(r2["load"]) > 0, strat
        assert r2["imbalance"] <= bound + 1e-6
    # contiguous strategy keeps forward order
    t3, _ = partition_tensors(params, num_parts=4, strategy="contiguous")
    owners = list(t3.values())
    assert owners == sorted(owners)


def test_validation_verbose_and_malloc(capsys):
    with pytest.raises(AssertionError):
        partition_tensors(_toy(), num_parts=2, evenness_priority=1.5)
    with pytest.raises(AssertionError):
        partition_tensors(_toy(), num_parts=0)
    with pytest.raises(AssertionError):
        partition_tensors(_toy(), num_parts=2, malloc=True)
    partition_tensors(_toy(), num_parts=2, verbose=True)   # the reference crashes here (no ranks_map)
    assert "partition t0" in capsys.readouterr().out
    table, tensors = partition_tensors(_toy(), ranks_map=["cpu", "cpu"], malloc=True)
    assert all(t.device.type == "cpu" for t in tensors.values())
