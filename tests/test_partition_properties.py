"""Property tests (hypothesis) of the three partition planners behind ``partition_tensors``."""
import itertools
from collections import OrderedDict

import torch
from hypothesis import given, settings, strategies as st

from tiny_deepspeed_b200.parallel.partition import (_assign_lpt, _split_contiguous, _walk_greedy, partition_tensors,
                                                     partition_report)

sizes_st = st.lists(st.integers(min_value=1, max_value=5000), min_size=1, max_size=9)


def _loads(sizes, assign, k):
    load = [0] * k
    for s, a in zip(sizes, assign):
        load[a] += s
    return load


def _best_contiguous(sizes, k):
    n, best = len(sizes), None
    for cuts in itertools.combinations(range(1, n), min(k - 1, n - 1)):
        bounds = (0,) + cuts + (n,)
        worst = max(sum(sizes[a:b]) for a, b in zip(bounds, bounds[1:]))
        best = worst if best is None else min(best, worst)
    return best if best is not None else sum(sizes)


@settings(max_examples=150, deadline=None)
@given(sizes_st, st.integers(min_value=1, max_value=5))
def test_contiguous_split_is_optimal_and_monotone(sizes, k):
    a = _split_contiguous(sizes, k)
    assert len(a) == len(sizes) and all(0 <= x < k for x in a)
    assert all(x <= y for x, y in zip(a, a[1:]))                       # contiguous: owners never decrease
    assert max(_loads(sizes, a, k)) == _best_contiguous(sizes, k)      # bottleneck-optimal among contiguous splits


def _opt_makespan(sizes, k):
    best = sum(sizes)
    for assign in itertools.product(range(k), repeat=len(sizes)):
        best = min(best, max(_loads(sizes, assign, k)))
    return best


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(min_value=1, max_value=5000), min_size=1, max_size=7), st.integers(min_value=1, max_value=4))
def test_lpt_respects_the_classic_bound(sizes, k):
    a = _assign_lpt(sizes, k)
    assert len(a) == len(sizes) and all(0 <= x < k for x in a)
    load = _loads(sizes, a, k)
    assert max(load) <= (4.0 / 3.0 - 1.0 / (3 * k)) * _opt_makespan(sizes, k) + 1e-9      # Graham's LPT guarantee
    if len(sizes) >= k:
        assert min(load) > 0                                           # nobody idles when there is enough work


@settings(max_examples=150, deadline=None)
@given(sizes_st, st.integers(min_value=1, max_value=5), st.floats(min_value=0.0, max_value=1.0))
def test_greedy_walk_is_forward_ordered_and_total(sizes, k, e):
    a = _walk_greedy(sizes, k, e)
    assert len(a) == len(sizes) and a[0] == 0 and all(0 <= x < k for x in a)
    assert all(0 <= y - x <= 1 for x, y in zip(a, a[1:]))              # reference semantics: walk forward, one part at a time


@settings(max_examples=50, deadline=None)
@given(sizes_st, st.integers(min_value=1, max_value=4), st.sampled_from(["greedy", "contiguous", "balanced"]))
def test_public_api_assigns_every_tensor_once(sizes, k, strategy):
    with torch.device("meta"):
        tensors = OrderedDict((f"t{i}", torch.empty(s)) for i, s in enumerate(sizes))
    table, same = partition_tensors(tensors, num_parts=k, strategy=strategy)
    assert same is tensors and list(table) == list(tensors)
    assert all(isinstance(v, int) and 0 <= v < k for v in table.values())
    rep = partition_report(tensors, table, k)
    assert sum(rep["load"]) == sum(sizes) and len(rep["load"]) == k
