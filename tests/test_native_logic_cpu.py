"""Host-side planning logic of the NVLink policy, exercised without GPUs: ZeRO-3 fetch-group construction and the owner
ranges of the post-step parameter multicast (tiny_deepspeed_b200/parallel/native_policy.py).  The kernels themselves are
covered by tests/test_gpu_comm.py."""
from collections import OrderedDict

import torch

from tiny_deepspeed_b200.parallel.native_policy import NativePolicy, _pad, ALIGN
from tiny_deepspeed_b200.parallel.partition import partition_tensors
from tiny_deepspeed_b200.models.gpt2 import GPT2Model, gpt2_config


def _stub(mode, world, esize=2, strategy="contiguous"):
    cfg = gpt2_config("tiny", n_layer=3, n_embd=128, n_head=4, vocab_size=1024, block_size=64)
    with torch.device("meta"):
        model = GPT2Model(cfg)
        named = OrderedDict(model.named_parameters())
        table, _ = partition_tensors(named, num_parts=world, strategy=strategy)
    pol = object.__new__(NativePolicy)
    pol.mode, pol.world, pol.rank, pol.table, pol.esize = mode, world, 0, table, esize
    pol.names = list(named)
    pol.numel = {n: p.numel() for n, p in named.items()}
    pol.poff, share, off = {}, [0] * world, 0
    for n in pol.names:
        if mode == "zero3":
            pol.poff[n] = share[table[n]]
            share[table[n]] += _pad(pol.numel[n])
        else:
            pol.poff[n] = off
            off += _pad(pol.numel[n])
    pol.slot_bytes = (max(_pad(v) for v in pol.numel.values()) * esize + 4095) // 4096 * 4096
    pol.comm_stub = True
    return pol, named


def test_pad_alignment():
    assert _pad(1) == ALIGN and _pad(ALIGN) == ALIGN and _pad(ALIGN + 1) == 2 * ALIGN


def test_zero3_fetch_groups_cover_sequence_and_respect_slots():
    for esize in (2, 4):
        pol, named = _stub("zero3", 4, esize)
        # forward use order then backward (reverse) order, as acquire() records them
        pol._seq = list(pol.names) + list(reversed(pol.names))
        pol._build_groups()
        assert sorted(pol._group_of) == list(range(len(pol._seq)))          # every use belongs to a group
        cap = min(pol.slot_bytes // esize, (16 << 20) // esize)
        for pos, n in enumerate(pol._seq):
            g = pol._groups[pol._group_of[pos]]
            assert g["owner"] == pol.table[n]                               # one owner per group
            assert g["lo"] <= pol.poff[n] and pol.poff[n] + _pad(pol.numel[n]) <= g["hi"]   # range contains the tensor
            assert (g["hi"] - g["lo"]) * esize <= pol.slot_bytes            # fits a staging slot
            assert g["hi"] - g["lo"] <= max(cap, _pad(pol.numel[n]))
        firsts = [g["first"] for g in pol._groups]
        assert firsts == sorted(firsts)                                     # groups are ordered by first use
        assert len(pol._groups) < len(pol._seq)                             # something was actually merged


def test_broadcast_ranges_partition_the_buffer_by_owner():
    pol, named = _stub("zero1", 3)
    pol.broadcast_params()                                                  # comm_stub: only plans the ranges
    ranges = pol._bcast_ranges
    covered = 0
    for (lo, hi, owner), nxt in zip(ranges, ranges[1:] + [None]):
        assert lo == covered and hi > lo
        covered = hi
        if nxt is not None:
            assert nxt[2] != owner or nxt[0] != hi                          # maximal merge of adjacent same-owner tensors
    assert covered == sum(_pad(v) for v in pol.numel.values())
    for n in pol.names:                                                     # every tensor lies inside a range of its owner
        r = next(r for r in ranges if r[0] <= pol.poff[n] < r[1])
        assert r[2] == pol.table[n] and pol.poff[n] + _pad(pol.numel[n]) <= r[1]
    # contiguous partition -> exactly one range per rank
    assert len(ranges) == 3


def test_broadcast_is_a_noop_for_ddp_and_zero3():
    for mode in ("ddp", "zero3"):
        pol, _ = _stub(mode, 2)
        pol.broadcast_params()
        assert getattr(pol, "_bcast_ranges", None) is None


# ---------------------------------------------------------------------------------------------------------------------
# round 2: bucketed ZeRO step inside backward + gradient ring.  The CUDA streams / events / kernels are replaced by
# recorders so the host-side schedule (which bucket is launched when, with which ranges, which slot waits) runs on CPU.
# ---------------------------------------------------------------------------------------------------------------------
class _FakeStream:
    def __init__(self, name, log):
        self.name, self.log = name, log

    def wait_stream(self, other):
        self.log.append(("wait_stream", self.name, other.name))

    def wait_event(self, ev):
        self.log.append(("wait_event", self.name, ev.tag))


class _FakeEvent:
    n = 0

    def __init__(self, *a, **k):
        _FakeEvent.n += 1
        self.tag = _FakeEvent.n
        self.on = None

    def record(self, stream=None):
        self.on = getattr(stream, "name", None)


class _FakeExt:
    COMM_MAX_RANGES = 320
    COMM_MAX_BLOCKS = 128

    def __init__(self, log):
        self.log = log

    def comm_zero_fused_adam(self, ctx, gbuf, pbuf, ranges, master, m, v, *hyper):
        self.log.append(("fused", [list(r) for r in ranges], hyper[-1]))
        return max(1, hyper[-1])

    def step_increment(self, t):
        t += 1


def _policy_for_schedule(monkeypatch, mode, world, rank, bucket_bytes, log):
    import contextlib
    from tiny_deepspeed_b200 import ops
    import tiny_deepspeed_b200 as tds
    cfg = gpt2_config("tiny", n_layer=3, n_embd=128, n_head=4, vocab_size=1024, block_size=64)
    with torch.device("meta"):
        meta = OrderedDict(GPT2Model(cfg).named_parameters())
        table, _ = partition_tensors(meta, num_parts=world, strategy="balanced")
    model = GPT2Model(cfg).to(torch.bfloat16)
    named = list(model.named_parameters())
    cur = _FakeStream("compute", log)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: cur)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    fake = _FakeExt(log)
    monkeypatch.setattr(ops, "ext", lambda: fake)
    pol = object.__new__(NativePolicy)
    pol.mode, pol.world, pol.rank, pol.table = mode, world, rank, table
    pol.device, pol.dtype, pol.f32, pol.esize, pol.scale = torch.device("cpu"), torch.bfloat16, False, 2, 1.0
    pol.names = [n for n, _ in named]
    pol.params = OrderedDict(named)
    pol.shape = {n: tuple(p.shape) for n, p in named}
    pol.numel = {n: p.numel() for n, p in named}
    pol._plan_gradient_layout(bucket_bytes, grad_accumulation=False, ring_slots=3)
    pol.poff, pol.ptotal = dict(pol.foff), pol.ftotal
    pol.gflat = torch.zeros(pol.gtotal, dtype=torch.bfloat16)
    pol.pflat = torch.zeros(pol.ptotal, dtype=torch.bfloat16)
    pol.gview = {n: pol.gflat[pol.goff[n]: pol.goff[n] + pol.numel[n]].view(pol.shape[n]) for n in pol.names}
    pol._name_of = {id(p): n for n, p in named}
    pol.comm_stream = _FakeStream("comm", log)
    pol.step_stream = pol.comm_stream
    pol.comm_stub, pol.fused_rs, pol.rs_names = True, False, frozenset()
    pol._solo_ctx = pol._solo_g = pol._solo_p = None
    pol._accumulated, pol._opt_state = set(), None
    pol.stats = {"allreduce_launches": 0, "fused_steps": 0, "bytes": 0}
    pol.fetch, pol._seq, pol._seq_frozen, pol._fetched, pol._pos = "push", [], False, {}, 0
    pol._reset_round()
    for n, p in named:
        p._tds_policy, p._tds_name, p.bwd_sync = pol, n, True
    opt = tds.AdamW(named, lr=1e-3)
    pol.bind_optimizer(opt)
    return pol, opt, named, table


def test_gradient_ring_layout_never_overlaps_live_buckets():
    log = []
    import pytest
    mp = pytest.MonkeyPatch()
    try:
        pol, opt, named, table = _policy_for_schedule(mp, "zero2", 4, 1, 64 << 10, log)
        assert pol.ring and pol.nslots_g == 3 and len(pol.buckets) > 4
        assert pol.gtotal == pol.nslots_g * pol.slot_elems < pol.ftotal          # smaller than the full gradient buffer
        for bi, b in enumerate(pol.buckets):
            lo = min(pol.goff[n] for n in b)
            hi = max(pol.goff[n] + _pad(pol.numel[n]) for n in b)
            slot = bi % pol.nslots_g
            assert slot * pol.slot_elems <= lo and hi <= (slot + 1) * pol.slot_elems      # a bucket lives in ONE slot
            spans = sorted((pol.goff[n], pol.goff[n] + _pad(pol.numel[n])) for n in b)
            assert all(a[1] <= c[0] for a, c in zip(spans, spans[1:]))                    # tensors do not overlap
        # zero1 keeps the full layout
        pol1, *_ = _policy_for_schedule(mp, "zero1", 4, 1, 64 << 10, [])
        assert not pol1.ring and pol1.goff == pol1.foff and pol1.gtotal == pol1.ftotal
    finally:
        mp.undo()


def test_zero_buckets_step_inside_backward_in_order_with_slot_waits():
    import pytest
    mp = pytest.MonkeyPatch()
    try:
        for mode in ("zero1", "zero2", "zero3"):
            log = []
            pol, opt, named, table = _policy_for_schedule(mp, mode, 4, 2, 64 << 10, log)
            nb = len(pol.buckets)
            launched_before_step = 0
            for n, p in reversed(named):                      # backward visits parameters in reverse registration order
                out, acc = pol.grad_out(p)
                assert out.data_ptr() == pol.gview[n].data_ptr() and not acc
                pol.grad_ready(p, out)
            launched_before_step = sum(1 for e in log if e[0] == "fused")
            assert 0 < launched_before_step < nb              # most buckets went during backward, the last ones wait for step()
            assert pol.fused_optimizer_step(opt) is True
            fused = [e for e in log if e[0] == "fused"]
            assert len(fused) == nb                           # every bucket exactly once
            owned = {n for n in pol.names if table[n] == pol.rank}
            seen = []
            for (_, ranges, min_launches), names in zip(fused, pol.buckets):      # launched in bucket order
                mine = sorted((n for n in names if n in owned), key=lambda k: pol.goff[k])
                assert [r[0] for r in ranges] == [pol.goff[n] for n in mine]
                assert [r[3] for r in ranges] == [pol.poff[n] for n in mine]
                assert min_launches >= 1
                seen += mine
            assert sorted(seen) == sorted(owned)              # every owned tensor updated exactly once
            assert opt.step_count == 1
            # the compute stream joined the communication stream at the end of the step
            assert ("wait_stream", "compute", "comm") in log
            if pol.ring:
                waits = [e for e in log if e[0] == "wait_event" and e[1] == "compute"]
                assert len(waits) == nb - pol.nslots_g        # one slot wait per recycled bucket
            # second round works from a clean state
            n_before = len([e for e in log if e[0] == "fused"])
            for n, p in reversed(named):
                p.bwd_sync = True
                out, _ = pol.grad_out(p)
                pol.grad_ready(p, out)
            assert pol.fused_optimizer_step(opt) is True
            assert len([e for e in log if e[0] == "fused"]) == 2 * n_before and opt.step_count == 2
    finally:
        mp.undo()


def test_bucket_runs_merge_adjacent_tensors_of_one_owner():
    import pytest
    mp = pytest.MonkeyPatch()
    try:
        pol, opt, named, table = _policy_for_schedule(mp, "zero2", 2, 0, 1 << 20, [])
        for b in range(len(pol.buckets)):
            runs = pol._bucket_runs(b)
            flat = [n for _, names in runs for n in names]
            assert sorted(flat) == sorted(pol.buckets[b])
            for (o1, n1), (o2, n2) in zip(runs, runs[1:]):
                adjacent = pol.goff[n1[-1]] + _pad(pol.numel[n1[-1]]) == pol.goff[n2[0]]
                assert o1 != o2 or not adjacent              # maximal merge
            for o, names in runs:
                assert all(table[n] == o for n in names)
    finally:
        mp.undo()


def test_flash_backward_work_split_covers_every_step_once():
    """Host-side model of flash_bwd_kernel's work mapping (csrc/flash_sm100.cu): key block jb meets the query blocks i >= jb;
    blocks with more than ceil(nq / 2) steps are cut into two CTAs.  Every (jb, i) pair must be covered exactly once, no CTA
    may be empty, and the longest CTA has ceil(nq / 2) steps (the point of the split)."""
    def cta(x, nq):
        n_split = nq - (nq + 1) // 2
        if x < 2 * n_split:
            jb = x >> 1
            n_all = nq - jb
            h0 = (n_all + 1) // 2
            it0, n_it = (h0, n_all - h0) if x & 1 else (0, h0)
            return jb, it0, n_it, True
        jb = x - n_split
        return jb, 0, nq - jb, False

    for nq in range(1, 33):
        n_split = nq - (nq + 1) // 2
        seen, longest = {}, 0
        for x in range(nq + n_split):
            jb, it0, n_it, split = cta(x, nq)
            assert 0 <= jb < nq and n_it >= 1, (nq, x)
            assert split == (jb < n_split)
            longest = max(longest, n_it)
            for it in range(it0, it0 + n_it):
                i = jb + it
                assert jb <= i < nq
                assert (jb, i) not in seen, (nq, jb, i)
                seen[(jb, i)] = x
        assert len(seen) == nq * (nq + 1) // 2
        assert longest == (nq + 1) // 2
        # the diagonal step (i == jb, the only masked one) always belongs to the CTA that starts at it0 == 0
        for jb in range(nq):
            x = seen[(jb, jb)]
            assert cta(x, nq)[1] == 0
