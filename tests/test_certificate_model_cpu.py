"""A numpy model of the "lists of 32 + shared bound + floor certificate" argument used by the brute-force path for 24 < k <= 64
(csrc/brute_force.cu: list_floor_kernel, exact.cu: approx_floor) and by the fused IVF coarse search (csrc/ivf_common.cu:
coarse_merge_kernel).  The model replays what the scan epilogue does — every list keeps its KC best entries, rejects an element
that is not strictly better than min(its own KC-th entry, the bound shared between the lists), and publishes its KC-th entry as
the new shared bound whenever it is full — with the lists visited in a random interleaving, and checks the invariant the
kernels rely on:

    every element outside the candidate set scores >= floor = min over the FULL lists of their worst kept entry,

hence: if the k-th best candidate is strictly below the floor, the k best candidates ARE the k best elements (ties included).
It also checks that the certificate does fail when it has to (more than KC of the true top-k inside one list)."""
import numpy as np
import pytest


def scan(scores, list_of, n_lists, KC, rng):
    """Returns (candidate element ids, floor).  Elements arrive in a random global order (= tiles of different work items
    interleaving on different SMs)."""
    kept = [[] for _ in range(n_lists)]  # (score, id), unsorted
    shared = np.inf
    for e in rng.permutation(len(scores)):
        l, s = list_of[e], scores[e]
        own = max(k for k, _ in kept[l]) if len(kept[l]) == KC else np.inf
        if not s < min(own, shared):
            continue
        if len(kept[l]) == KC:
            kept[l].remove(max(kept[l]))
        kept[l].append((s, e))
        if len(kept[l]) == KC:
            shared = min(shared, max(k for k, _ in kept[l]))  # atomicMin of the list's KC-th entry
    floor = min([max(k for k, _ in lst) for lst in kept if len(lst) == KC], default=np.inf)
    return np.array([e for lst in kept for _, e in lst], dtype=np.int64), floor


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("k", [10, 32, 48, 64])
def test_floor_bounds_everything_outside_the_candidates_and_the_certificate_is_sound(seed, k):
    rng = np.random.default_rng(seed)
    n, n_lists, KC = 4000, int(rng.integers(2, 17)), 32
    # ties on purpose: scores drawn from a coarse grid in half of the runs
    scores = rng.standard_normal(n).astype(np.float32) if seed % 2 else np.round(rng.standard_normal(n), 1).astype(np.float32)
    list_of = rng.integers(0, n_lists, n)
    if seed % 3 == 0:  # adversarial: the best elements crowd into list 0
        list_of[np.argsort(scores)[: 3 * KC]] = 0
    cand, floor = scan(scores, list_of, n_lists, KC, rng)
    outside = np.setdiff1d(np.arange(n), cand)
    if outside.size:
        assert scores[outside].min() >= floor, "an element outside the candidates beats the floor"
    order = np.lexsort((np.arange(n), scores))          # (score, id) order = select_k's tie rule on a dense row
    c_order = cand[np.lexsort((cand, scores[cand]))]
    if len(c_order) >= k and scores[c_order[k - 1]] < floor:
        np.testing.assert_array_equal(c_order[:k], order[:k])  # certified => identical selection, ties included
    else:
        # not certified: the kernels fall back to the dense / exact path; nothing to check except that this only happens when
        # a list really could have dropped a top-k element (some full list's worst entry is within the top-k score range)
        assert floor <= scores[order[k - 1]] or len(c_order) < k


def test_the_certificate_fails_when_one_list_holds_more_than_KC_of_the_top_k():
    rng = np.random.default_rng(0)
    n, KC, k = 2000, 32, 48
    scores = rng.standard_normal(n).astype(np.float32)
    list_of = rng.integers(1, 8, n)
    list_of[np.argsort(scores)[:40]] = 0   # 40 of the true top-48 in ONE list of 32
    cand, floor = scan(scores, list_of, 8, KC, rng)
    c_sorted = np.sort(scores[cand])
    assert not (len(cand) >= k and c_sorted[k - 1] < floor)
