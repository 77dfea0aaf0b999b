"""The tie-break of equal weights comes from Rust std's BinaryHeap (push = sift up while strictly lighter; pop = move the
last element to the root, sift the hole DOWN TO THE BOTTOM taking the right child unless the left one is strictly lighter,
then sift up), which is outside /root/reference and cannot be run here ("parity unpinned").

CPython's `heapq` is the same algorithm written by other hands (heappop -> _siftup walks the hole to a leaf choosing the right
child unless `left < right`, then _siftdown; heappush -> _siftdown moves up only while `new < parent`).  Building the trees
with it is therefore an implementation of Appendix A.2 that none of this repository's three restatements (product C++, C
oracle, Python restatement) shares a line with: if they all agree with it on tie-heavy inputs, the restatements implement
that algorithm faithfully.  (That std's source IS that algorithm remains the stated assumption.)"""
import heapq

import numpy as np
import pytest

from huff_encoding_b200.api import HuffTree
from oracle import oracle as O


class _Item:
    """compares by weight only (branch_heap.rs:67-83: a HuffBranch orders by its leaf's weight, reversed for a min-heap)"""
    __slots__ = ("weight", "letter", "left", "right")

    def __init__(self, weight, letter=None, left=None, right=None):
        self.weight, self.letter, self.left, self.right = weight, letter, left, right

    def __lt__(self, other):
        return self.weight < other.weight


def heapq_codes(pairs):
    heap = []
    for letter, w in pairs:                                  # branch_heap.rs:52-58: sequential pushes, not heapify
        heapq.heappush(heap, _Item(w, letter))
    while len(heap) > 1:                                     # tree_inner.rs:289-303
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        heapq.heappush(heap, _Item(a.weight + b.weight, None, a, b))
    root = heapq.heappop(heap)
    codes = {}
    if root.left is None:
        return {root.letter: "0"}
    stack = [(root, "")]
    while stack:                                             # left = 0, right = 1; DFS left first, later visit wins
        node, code = stack.pop()
        if node.left is None:
            codes[node.letter] = code
        else:
            stack.append((node.right, code + "1"))
            stack.append((node.left, code + "0"))
    return codes


@pytest.mark.parametrize("seed", range(40))
def test_heapq_built_trees_match_the_oracle_and_the_product_on_tie_heavy_weights(seed):
    rng = np.random.default_rng(31337 + seed)
    n = int(rng.choice([2, 3, 4, 5, 7, 8, 16, 33, 64, 100, 200, 256]))
    letters = sorted(int(x) for x in rng.choice(256, size=n, replace=False))
    kind = seed % 4
    if kind == 0:
        w = np.ones(n, dtype=np.int64)                                        # everything ties
    elif kind == 1:
        w = rng.integers(1, 4, size=n)                                        # three weight classes
    elif kind == 2:
        w = np.sort(rng.integers(1, 6, size=n)) * (1 << rng.integers(0, 3, size=n))   # sums that collide with leaves
    else:
        w = rng.integers(1, 50, size=n)
    pairs = [(l, int(x)) for l, x in zip(letters, w)]
    want = heapq_codes(pairs)
    assert O.tree_from_pairs([p[0] for p in pairs], [p[1] for p in pairs]).codes() == want
    assert HuffTree.from_weights(pairs).read_codes() == want


def test_heapq_trees_on_the_reference_goldens():
    # tests/tree_init.rs:10-47 and the abbccc doctest
    assert heapq_codes([(0, 5), (1, 9), (2, 12), (3, 13), (4, 16), (5, 45)]) == \
        {0: "1100", 1: "1101", 2: "100", 3: "101", 4: "111", 5: "0"}
    assert heapq_codes([(ord("a"), 1), (ord("b"), 2), (ord("c"), 3)]) == {ord("c"): "0", ord("a"): "10", ord("b"): "11"}
