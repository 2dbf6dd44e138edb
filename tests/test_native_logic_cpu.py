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
