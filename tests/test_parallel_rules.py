"""The two index rules the GPU BVH build and the device-side flatten rest on (csrc/bvh_build_gpu.cu, csrc/flatten_gpu.cu),
checked on the CPU against the sequential definitions they replace.

1. The reference's in-place partition (cpu/src/bvh.c:244-259), `for i: if left(A[i]) swap(A[i], A[n_left++])`, equals: lefts
   compacted in encounter order; the right element that starts at x ends at the first element of the chain
   x -> posL[x] -> posL[posL[x]] ... that is >= n_left, where posL[r] is the position of the left of rank r.
2. In the reference's node numbering the k-th inner node in depth-first pre-order has its children at 1 + 2k.
(The wide trees of the fast build are built level by level from shared host/device code, csrc/wide8.h; their rules are checked by
tests/test_flatten_host.py and tests/test_wide8_host.py.)
"""
import numpy as np
import pytest

from conftest import GOLD


def forward_swap_partition(a, left):
    a = list(a)
    nl = 0
    for i in range(len(a)):
        if left[a[i]]:
            a[i], a[nl] = a[nl], a[i]
            nl += 1
    return a, nl


def chain_partition(a, left):
    n = len(a)
    flags = [left[v] for v in a]
    pos_l = [i for i in range(n) if flags[i]]
    nl = len(pos_l)
    out = [None] * n
    for r, i in enumerate(pos_l):
        out[r] = a[i]
    longest = 0
    for x in range(n):
        if flags[x]:
            continue
        p, steps = x, 0
        while p < nl:
            p = pos_l[p]
            steps += 1
        assert out[p] is None
        out[p] = a[x]
        longest = max(longest, steps)
    return out, nl, longest


@pytest.mark.parametrize("n", [1, 2, 3, 8, 33, 257, 2000])
@pytest.mark.parametrize("p_left", [0.0, 0.02, 0.5, 0.98, 1.0])
def test_chain_rule_equals_the_forward_swap_partition(n, p_left):
    rng = np.random.default_rng(n * 131 + int(p_left * 100))
    for _ in range(5):
        a = rng.permutation(n).tolist()
        left = (rng.random(n) < p_left).tolist()
        want, nl = forward_swap_partition(a, left)
        got, nl2, _ = chain_partition(a, left)
        assert nl == nl2 and got == want


def test_chain_rule_on_adversarial_patterns():
    n = 500
    a = list(range(n))
    for name, left in {
        "one right, then lefts (one chain of length n-1)": [False] + [True] * (n - 1),
        "rights first": [False] * (n // 2) + [True] * (n - n // 2),
        "lefts first": [True] * (n // 2) + [False] * (n - n // 2),
        "alternating": [i % 2 == 0 for i in range(n)],
        "alternating, right first": [i % 2 == 1 for i in range(n)],
        "runs": [(i // 7) % 2 == 0 for i in range(n)],
    }.items():
        want, nl = forward_swap_partition(a, left)
        got, nl2, longest = chain_partition(a, left)
        assert (got, nl2) == (want, nl), name
        total = sum(1 for v in left if v)
        assert longest <= max(total, 1)
    # the total length of all chains is bounded by the number of lefts (each left moves one right)
    left = [False] + [True] * (n - 1)
    assert chain_partition(a, left)[2] == n - 1


@pytest.mark.parametrize("scene", ["car_only", "car_boxed", "soup2k"])
def test_record_index_rules_of_the_device_side_flatten(rt, scene):
    sc = rt.Scene.load_rtsc(GOLD / "scenes" / f"{scene}.rtsc").build_bvh(6)
    arr = sc.arrays()
    sc.close()
    nodes = np.frombuffer(arr["bvh_nodes"].tobytes(), dtype=[("mn", "<f4", 3), ("mx", "<f4", 3), ("tr_len", "<i4"), ("idx", "<i4")])
    inner = (nodes["tr_len"] == 0) & (nodes["idx"] != 0)
    assert len(nodes) == 1 + 2 * int(inner.sum())  # every split allocated two nodes
    # depth-first pre-order over the inner nodes, left subtree first (flatten.cpp pass 1)
    order, depth_of, stack = [], {}, [(0, 0)] if inner[0] else []
    while stack:
        v, d = stack.pop()
        order.append(v)
        depth_of[v] = d
        c = int(nodes["idx"][v])
        if inner[c + 1]:
            stack.append((c + 1, d + 1))
        if inner[c]:
            stack.append((c, d + 1))
    assert len(order) == int(inner.sum())
    for k, v in enumerate(order):
        assert (int(nodes["idx"][v]) - 1) // 2 == k
