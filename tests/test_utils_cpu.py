"""Auxiliary subsystems on CPU: checkpoint round trip (single process), watchdog, metrics logger, step timer, autotuner."""
import json
import time

import torch

import tiny_deepspeed_b200 as tds
from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config
from tiny_deepspeed_b200.utils import (save_checkpoint, load_checkpoint, Watchdog, MetricsLogger, StepTimer,
                                       format_loss_line, nvtx_range)


def _tiny():
    torch.manual_seed(0)
    cfg = gpt2_config("tiny", n_layer=1, n_embd=32, n_head=2, vocab_size=64, block_size=16)
    return cfg, GPT2Model(cfg)


def test_checkpoint_roundtrip_single_process(tmp_path):
    cfg, m = _tiny()
    opt = tds.AdamW(m.named_parameters(), lr=1e-2)
    x = torch.randint(0, cfg.vocab_size, (2, 16)); y = torch.randint(0, cfg.vocab_size, (2, 16))
    for _ in range(2):
        _, l = m(x, y); l.backward(); opt.step()
    path = save_checkpoint(str(tmp_path), m, opt, step=2, extra={"note": "hi"})
    assert path.endswith("shard_00000_of_00001.pt")
    cfg2, m2 = _tiny()
    with torch.no_grad():
        for p in m2.parameters():
            p.add_(1.0)
    opt2 = tds.AdamW(m2.named_parameters(), lr=1e-2)
    meta = load_checkpoint(str(tmp_path), m2, opt2)
    assert meta["step"] == 2 and meta["extra"]["note"] == "hi" and opt2.step_count == 2
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.equal(a, b), n
    # resumed training continues identically
    _, l1 = m(x, y); l1.backward(); opt.step()
    _, l2 = m2(x, y); l2.backward(); opt2.step()
    assert torch.equal(l1, l2)
    for a, b in zip(m.parameters(), m2.parameters()):
        torch.testing.assert_close(a, b)


def test_watchdog_fires_and_disarms():
    wd = Watchdog(timeout_s=0.3, name="unit", abort=False)
    with wd:
        time.sleep(0.05)
    assert not wd.fired
    wd.arm()
    time.sleep(0.9)
    assert wd.fired
    wd.close()


def test_metrics_logger_and_helpers(tmp_path):
    p = tmp_path / "m.jsonl"
    ml = MetricsLogger(str(p))
    ml.log(step=1, loss=2.5)
    rec = json.loads(p.read_text().splitlines()[0])
    assert rec["step"] == 1 and rec["loss"] == 2.5 and "t" in rec
    assert format_loss_line(3, 1.23456) == "iter 3 loss: 1.2346"      # byte-compatible with the reference print
    t = StepTimer("cpu")
    t.start(); time.sleep(0.01); ms = t.stop()
    assert ms >= 5 and t.mean_ms() == ms
    with nvtx_range("noop"):
        pass


def test_trainstep_rejects_shape_change_only_when_captured():
    cfg, m = _tiny()
    opt = tds.SGD(m.named_parameters(), lr=0.1, momentum=0.9)
    step = tds.TrainStep(m, opt)          # CPU: eager, any shape goes
    for T in (8, 16):
        x = torch.randint(0, cfg.vocab_size, (1, T))
        assert torch.isfinite(step(x, x))
    assert opt.step_count == 2


def test_trainstep_wires_metrics_and_watchdog(tmp_path):
    """the aux subsystems are driven by the engine, not shelfware."""
    import json
    import torch
    import tiny_deepspeed_b200 as tds
    from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config
    from tiny_deepspeed_b200.utils import check_device_flags

    torch.manual_seed(0)
    cfg = gpt2_config("tiny")
    model = GPT2Model(cfg)
    opt = tds.AdamW(model.named_parameters(), lr=1e-3)
    path = tmp_path / "metrics.jsonl"
    step = tds.TrainStep(model, opt, use_graph=False, watchdog_s=60.0, metrics_path=str(path), metrics_every=2)
    assert step.watchdog is not None and step.metrics is not None
    x = torch.randint(0, cfg.vocab_size, (2, 16))
    for _ in range(5):
        loss = step(x, x)
    step.finish()
    assert not step.watchdog.fired
    step.watchdog.close()
    rows = [json.loads(l) for l in open(path)]
    assert [r["step"] for r in rows] == [2, 4] and all("loss" in r for r in rows)
    assert abs(rows[-1]["loss"] - float(loss)) < 10.0

    class _Comm:
        error = torch.zeros(1, dtype=torch.int32)

    class _Pol:
        comm = _Comm()

    check_device_flags(_Pol())                      # clean flag: no exception
    _Comm.error = torch.ones(1, dtype=torch.int32)
    import pytest
    with pytest.raises(RuntimeError, match="timed out"):
        check_device_flags(_Pol())
    check_device_flags(None)                        # no policy: nothing to check


def test_optimizer_step_count_roundtrip_matches_device_counter():
    """state_dict() must carry the true step even when only the device-side counter advanced."""
    import torch
    import tiny_deepspeed_b200 as tds
    lin = torch.nn.Linear(4, 4)
    opt = tds.AdamW(lin.named_parameters(), lr=1e-2)
    for _ in range(3):
        lin(torch.randn(2, 4)).sum().backward()
        opt.step()
    sd = opt.state_dict()
    assert sd["step"] == 3
    # emulate graph replays: the device counter is ahead of the Python one
    opt._step_dev = torch.tensor([7], dtype=torch.int32)
    assert opt.state_dict()["step"] == 7 and opt.step_count == 7
    opt2 = tds.AdamW(lin.named_parameters(), lr=1e-2)
    opt2._step_dev = torch.tensor([0], dtype=torch.int32)
    opt2.load_state_dict(sd | {"step": 7})
    assert opt2.step_count == 7 and int(opt2._step_dev.item()) == 7


def test_prefetch_chain_registers_nothing_and_follows_forward_order():
    """The L2 prefetch chain (nn.modules.link_prefetch_chain) stores module references outside nn.Module's registries: parameter
    names / order — the compatibility contract with the reference's model — and the state dict are unchanged."""
    from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config
    model = GPT2Model(gpt2_config("tiny"))
    names = [n for n, _ in model.named_parameters()]
    assert len(names) == len(set(names))
    assert not any("_pf_" in n for n in names)
    assert not any("_pf_" in k for k in model.state_dict())
    blk0, blk1 = model.transformer.h[0], model.transformer.h[1]
    chain = [blk0.attn.c_attn, blk0.attn.c_proj, blk0.mlp.c_fc, blk0.mlp.c_proj, blk1.attn.c_attn]
    for a, b in zip(chain, chain[1:]):
        assert a.__dict__["_pf_fwd"] is b and b.__dict__["_pf_bwd"] is a
    last = model.transformer.h[-1].mlp.c_proj
    assert last.__dict__["_pf_fwd"] is model.lm_head and model.lm_head.__dict__["_pf_bwd"] is last
    assert "_pf_bwd" not in chain[0].__dict__ and "_pf_fwd" not in model.lm_head.__dict__
    # number of registered submodules is what the architecture defines (nothing sneaked in through the chain)
    n_linear = sum(1 for m in model.modules() if m.__class__.__name__ == "Linear")
    assert n_linear == 4 * len(model.transformer.h) + 1


def test_gpu_only_switches_are_noops_on_cpu():
    """ops.prefetch_next / ops.set_pdl must not touch the extension on a CPU box (the CPU suite runs without a GPU)."""
    import torch
    from tiny_deepspeed_b200 import ops
    ops.prefetch_next(torch.zeros(64))
    ops.prefetch_next(None)
    ops.set_pdl(False)
    ops.set_pdl(True, force=True)
