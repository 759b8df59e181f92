"""Design note as a test (no GPU): the (best, second best) pair the sorted grid scan keeps per landmark can be built from
ANY partition of the candidate scores — per-lane partial pairs merged in any order — which is what the warp-cooperative
rounds of assoc_grid_kernel rely on (assoc_scan.cu: score_staged, the xor-shuffle merge, the merge with the big-list pair).
The update and the merge are the kernel's own formulas:
    update:  b2 = min(b2, max(b1, s));  b1 = min(b1, s)
    merge:   b2 = min(max(b1, o1), min(b2, o2));  b1 = min(b1, o1)
Ties count twice (two candidates with the same best score give b2 == b1), so the certificate (strict gap between the two)
sends every tie to the exact re-rank: which of two equal scores names the winner's id never reaches the output."""
import numpy as np
from hypothesis import given, settings, strategies as st

INF = np.float32(np.inf)


def fold(scores):
    b1, b2 = INF, INF
    for s in scores:
        s = np.float32(s)
        b2 = min(b2, max(b1, s))
        b1 = min(b1, s)
    return b1, b2


def merge(a, o):
    return min(a[0], o[0]), min(max(a[0], o[0]), min(a[1], o[1]))


@settings(max_examples=300, deadline=None)
@given(st.lists(st.floats(min_value=0, max_value=1e6, width=32) | st.just(float("inf")), min_size=0, max_size=70),
       st.integers(min_value=1, max_value=16), st.randoms(use_true_random=False))
def test_partial_pairs_merge_to_the_two_smallest_scores(scores, parts, rnd):
    want = sorted(np.float32(s) for s in scores)[:2] + [INF, INF]
    want = (want[0], want[1])
    assert fold(scores) == want
    # deal the scores over `parts` lanes in strides (member slot x record index, as the cooperative round does), fold per
    # lane, merge the lane pairs in a random order
    lanes = [fold(scores[p::parts]) for p in range(parts)]
    rnd.shuffle(lanes)
    acc = (INF, INF)
    for pr in lanes:
        acc = merge(acc, pr)
    assert acc == want
    # butterfly order of the xor shuffles (parts padded to a power of two with empty lanes)
    n = 1
    while n < parts:
        n *= 2
    lanes = [fold(scores[p::parts]) for p in range(parts)] + [(INF, INF)] * (n - parts)
    off = 1
    while off < n:
        lanes = [merge(lanes[i], lanes[i ^ off]) for i in range(n)]
        off *= 2
    assert all(pr == want for pr in lanes)


def test_a_tie_for_the_best_score_leaves_no_gap():
    b1, b2 = fold([4.0, 2.5, 2.5, 9.0])
    assert b1 == b2 == np.float32(2.5)
    assert merge(fold([2.5, 9.0]), fold([4.0, 2.5])) == (np.float32(2.5), np.float32(2.5))
