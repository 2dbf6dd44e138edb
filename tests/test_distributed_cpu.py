"""Control flow, ownership, flag semantics and numerics of the four parallel modes on a gloo CPU process group
(the "fake backend" of SURVEY §4 item 4).  World size 2 and 3."""
from collections import OrderedDict

import pytest
import torch

from dist_utils import run_distributed


def _build(mode, rank, world, meta_zero3=True, strategy="greedy"):
    import tiny_deepspeed_b200 as tds
    from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config
    from tiny_deepspeed_b200.parallel import materialize_
    cfg = gpt2_config("tiny", n_layer=2, n_embd=32, n_head=2, vocab_size=64, block_size=16)
    with torch.device("meta"):
        meta = GPT2Model(cfg)
        parts, _ = tds.partition_tensors(OrderedDict(meta.named_parameters()), num_parts=world, strategy=strategy)
    if mode == "zero3" and meta_zero3:
        with torch.device("meta"):
            model = GPT2Model(cfg)
        model = tds.Zero3(model, parts, device="cpu", init_seed=7)
    else:
        with torch.device("meta"):
            model = GPT2Model(cfg)
        materialize_(model, device="cpu", seed=7)            # same deterministic init as the ZeRO-3 meta path
        W = {"ddp": tds.DDP, "zero1": tds.Zero1, "zero2": tds.Zero2, "zero3": tds.Zero3}[mode]
        model = W(model) if mode == "ddp" else W(model, parts)
    O = {"ddp": tds.DDPAdamW, "zero1": tds.Zero1AdamW, "zero2": tds.Zero2AdamW, "zero3": tds.Zero3AdamW}[mode]
    if mode == "ddp":
        opt = O(model.named_parameters(), lr=1e-2, weight_decay=0.1)
    else:
        opt = O(model.module.named_parameters(), lr=1e-2, weight_decay=0.1, param_part_table=parts,
                ranks_map=[f"cpu:{i}" for i in range(world)])
    return cfg, model, opt, parts


def _train(rank, world, mode, steps):
    import torch.distributed as dist
    cfg, model, opt, parts = _build(mode, rank, world)
    g = torch.Generator().manual_seed(100 + rank)      # different data per rank
    x = torch.randint(0, cfg.vocab_size, (2, 16), generator=g)
    y = torch.randint(0, cfg.vocab_size, (2, 16), generator=g)
    losses, checks = [], {}
    for i in range(steps):
        model.require_backward_grad_sync = True
        _, loss = model(x, y)
        assert model.require_backward_grad_sync is False      # one-shot flag consumed by forward (Q5)
        loss.backward()
        model.finish_grad_sync()
        if i == 0:
            named = dict(model.module.named_parameters())
            if mode in ("zero2", "zero3"):
                checks["nonowner_grad_none"] = all((p.grad is None and getattr(p, "_tds_grad", None) is None)
                                                   for n, p in named.items() if parts[n] != rank)
                checks["owner_grad_set"] = all(p.grad is not None for n, p in named.items() if parts[n] == rank)
            if mode == "zero3":
                checks["nonowner_param_freed"] = all(p.numel() == 0 for n, p in named.items() if parts[n] != rank)
            checks["bwd_sync_consumed"] = all(not p.bwd_sync for p in named.values())
            if mode != "ddp":
                checks["state_only_owned"] = None  # filled after step
        opt.step()
        if i == 0 and mode != "ddp":
            checks["state_only_owned"] = set(opt.state.keys()) == {n for n in parts if parts[n] == rank}
        l = loss.detach().clone()
        dist.all_reduce(l)
        losses.append(float(l) / world)
    # gather final full weights (owner holds truth for zero3)
    final = {}
    for n, p in model.module.named_parameters():
        shape = p._tds_shape
        if mode == "zero3":
            t = p.detach().clone() if parts[n] == rank else torch.empty(shape)
            dist.broadcast(t, src=parts[n])
        else:
            t = p.detach().clone()
        final[n] = t
    return losses, final, checks


@pytest.mark.parametrize("world", [2, 3])
def test_mode_equivalence(world):
    ref_losses, ref_final = None, None
    for mode in ("ddp", "zero1", "zero2", "zero3"):
        res = run_distributed(_train, world=world, args=(mode, 4))
        losses, final, checks = res[0]
        for r in range(world):
            assert all(v for v in res[r][2].values()), (mode, r, res[r][2])
            for n in final:                                   # replicas agree
                torch.testing.assert_close(res[r][1][n], final[n], rtol=0, atol=0, msg=f"{mode} {n}")
        assert losses[-1] < losses[0]
        if ref_losses is None:
            ref_losses, ref_final = losses, final
        else:
            assert losses == pytest.approx(ref_losses, rel=1e-5, abs=1e-6), mode
            for n in final:
                torch.testing.assert_close(final[n], ref_final[n], rtol=1e-4, atol=1e-6, msg=f"{mode} {n}")


def _accum(rank, world, mode):
    """Gradient accumulation: 2 micro-batches, sync only on the second -> same as one synced big step."""
    import torch.distributed as dist
    cfg, model, opt, parts = _build(mode, rank, world, meta_zero3=False)
    g = torch.Generator().manual_seed(5 + rank)
    xs = [torch.randint(0, cfg.vocab_size, (1, 16), generator=g) for _ in range(4)]
    model.require_backward_grad_sync = False
    _, l = model(xs[0], xs[1]); l.backward()
    model.require_backward_grad_sync = True
    _, l = model(xs[2], xs[3]); l.backward()
    model.finish_grad_sync()
    named = dict(model.module.named_parameters())
    name = "transformer.h.0.attn.c_attn.weight"
    owner = 0 if mode == "ddp" else parts[name]
    got = named[name].grad.clone() if rank == owner else torch.zeros(named[name]._tds_shape)
    dist.broadcast(got, src=owner)
    return got


def test_grad_accumulation_reduces_accumulated_sum():
    outs = {}
    for mode in ("ddp", "zero1"):
        outs[mode] = run_distributed(_accum, world=2, args=(mode,))[0]
    torch.testing.assert_close(outs["ddp"], outs["zero1"], rtol=1e-5, atol=1e-7)
    assert outs["ddp"].abs().sum() > 0


def _ckpt(rank, world, tmp):
    import tiny_deepspeed_b200 as tds
    from tiny_deepspeed_b200.utils import save_checkpoint, load_checkpoint
    cfg, model, opt, parts = _build("zero2", rank, world)
    x = torch.randint(0, cfg.vocab_size, (1, 16)); y = torch.randint(0, cfg.vocab_size, (1, 16))
    for _ in range(2):
        model.require_backward_grad_sync = True
        _, l = model(x, y); l.backward(); opt.step()
    save_checkpoint(tmp, model, opt, table=parts, step=2)
    before = {n: p.detach().clone() for n, p in model.module.named_parameters()}
    m_before = {n: st["exp_avg"].clone() for n, st in opt.state.items()}
    cfg2, model2, opt2, _ = _build("zero2", rank, world)
    meta = load_checkpoint(tmp, model2, opt2)
    ok = meta["step"] == 2 and opt2.step_count == 2
    for n, p in model2.module.named_parameters():
        ok = ok and torch.equal(p, before[n])
    for n in m_before:
        ok = ok and torch.equal(opt2.state[n]["exp_avg"], m_before[n])
    return ok


def test_sharded_checkpoint_roundtrip(tmp_path):
    assert all(run_distributed(_ckpt, world=2, args=(str(tmp_path),)))


def _owner_table_autobuild(rank, world):
    import tiny_deepspeed_b200 as tds
    from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config
    cfg = gpt2_config("tiny", n_layer=1, n_embd=16, n_head=2, vocab_size=32, block_size=8)
    torch.manual_seed(0)
    model = GPT2Model(cfg)
    with torch.device("meta"):
        parts, _ = tds.partition_tensors(OrderedDict(GPT2Model(cfg).named_parameters()), num_parts=world)
    model = tds.Zero1(model, parts)
    opt = tds.Zero1SGD(model.module.named_parameters(), lr=0.1, ranks_map=["cpu:0", "cpu:1"])   # no table given
    return opt.param_part_table == parts


def test_optimizer_builds_its_own_table():
    assert all(run_distributed(_owner_table_autobuild, world=2))
