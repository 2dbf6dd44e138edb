"""Public surface parity with the reference."""
import torch

EXPORTS = ["SGD", "AdamW", "DDPSGD", "DDPAdamW", "DDP", "Zero1SGD", "Zero1AdamW", "Zero1", "Zero2SGD",
           "Zero2AdamW", "Zero2", "Zero3SGD", "Zero3AdamW", "Zero3", "partition_tensors"]


def test_core_exports():
    import tiny_deepspeed_b200.core as core
    assert sorted(core.__all__) == sorted(EXPORTS)
    for n in EXPORTS:
        assert hasattr(core, n)


def test_reference_import_paths():
    from tiny_deepspeed.core import DDP, Zero1AdamW, partition_tensors  # noqa: F401
    from tiny_deepspeed.core.optim import SGD, AdamW  # noqa: F401
    from tiny_deepspeed.core.module import Linear, LayerNorm, Embedding  # noqa: F401
    from tiny_deepspeed.core.module.ops import (linear_forward, linear_input_grad, linear_weight_grad,  # noqa: F401
                                                linear_bias_grad, layernorm_fwd, layernorm_dx, layernorm_dwdb,
                                                embedding_forward, embedding_weight_grad)
    from tiny_deepspeed.core.autotuner import RuntimeAutoTuner  # noqa: F401
    from tiny_deepspeed.core.zero.utils.partition import partition_tensors as p2  # noqa: F401
    from tiny_deepspeed.core.zero.zero1 import Zero1, AdamW as Z1AdamW, Parameter  # noqa: F401
    from example.model import GPTConfig, GPT2Model  # noqa: F401
    import tiny_deepspeed_b200 as tds
    assert DDP is tds.DDP and Z1AdamW is tds.Zero1AdamW


def test_gpt2_parameter_order_and_counts():
    from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config
    expect = {"small": (101, 163.0), "medium": (197, 406.2), "large": (293, 838.1), "xl": (389, 1637.5)}
    for name, (ntens, mparams) in expect.items():
        with torch.device("meta"):
            m = GPT2Model(gpt2_config(name))
        names = [n for n, _ in m.named_parameters()]
        assert len(names) == ntens
        assert names[0] == "transformer.wte.weight" and names[1] == "transformer.wpe.weight"
        assert names[2:10] == [f"transformer.h.0.{s}" for s in
                               ("ln_1.weight", "ln_1.bias", "attn.c_attn.weight", "attn.c_proj.weight", "ln_2.weight",
                                "ln_2.bias", "mlp.c_fc.weight", "mlp.c_proj.weight")]
        assert names[-3:] == ["transformer.ln_f.weight", "transformer.ln_f.bias", "lm_head.weight"]
        assert abs(sum(p.numel() for p in m.parameters()) / 1e6 - mparams) < 0.06


def test_autotuner_keys_per_op():
    from tiny_deepspeed_b200.autotuner import RuntimeAutoTuner
    calls = []

    def slow(x):
        calls.append("slow")
        s = 0
        for _ in range(2000):
            s += 1
        return x + 1

    def fast(x):
        calls.append("fast")
        return x + 1

    t = RuntimeAutoTuner(enable=True, warmup_iterations=1, measure_iterations=5)
    x = torch.zeros(4)
    t.choose_function([slow, fast], x, key="fwd")
    assert t.best("fwd", x) == 1
    # a different op key is tuned independently (the reference caches ONE winner per tuner)
    t.choose_function([fast, slow], x, key="dx")
    assert t.best("dx", x) == 0
    t.final_tune()
    calls.clear()
    t.choose_function([slow, fast], torch.zeros(8), key="unseen")
    assert calls == ["slow"]  # finalized: unseen keys run candidate 0 without measuring


def test_inventory_parity_helpers():
    """Small pieces of the reference inventory: dtype table, conv placeholders,
    get_init_args, free-function collectives (single process: no-ops that keep semantics)."""
    import pytest
    import torch.nn as nn
    from tiny_deepspeed_b200.ops.utils import supported_acc_dtypes, acc_dtype
    from tiny_deepspeed_b200.nn.conv import Conv2d
    from tiny_deepspeed_b200.parallel import get_init_args, sync_grad, desync_grad, sync_param, desync_param_data
    assert supported_acc_dtypes[torch.bfloat16] is torch.float32 and acc_dtype(torch.int8) is torch.int32
    with pytest.raises(NotImplementedError):
        Conv2d(1, 1, 1)
    a = get_init_args(nn.Linear(3, 5, bias=False))
    assert a["in_features"] == 3 and a["out_features"] == 5 and a["bias"] is False
    assert get_init_args(nn.Embedding(7, 2))["num_embeddings"] == 7
    g = torch.ones(4)
    assert sync_grad(g) is None and desync_grad(g, 0) is g and desync_grad(g, 1) is None
    p = nn.Parameter(torch.ones(2, 3))
    full, h = sync_param(p, rank_id=0)
    assert full is p and h is None
    desync_param_data(p, rank_id=1)
    assert p.numel() == 0 and p._tds_shape == (2, 3)
