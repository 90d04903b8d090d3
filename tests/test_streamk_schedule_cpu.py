"""CPU: the stream-K schedule of the batched path iteration (csrc/gram_kernels.cu: path_step_sk_kernel)
restated in Python -- range per CTA, segments, workspace slots, contributors per tile -- and checked for
every property the kernel relies on, over many (tiles, k-steps, CTAs) combinations."""
import itertools

import pytest


def sk_begin(T, KT, P, c):
    return (T * KT * c) // P


def segments(T, KT, P, c):
    """(tile, k_lo, k_hi, slot or None) for CTA c, in processing order; slot None = whole tile."""
    r0, r1 = sk_begin(T, KT, P, c), sk_begin(T, KT, P, c + 1)
    out = []
    r = r0
    while r < r1:
        t = r // KT
        k_lo = r - t * KT
        k_hi = min(KT, r1 - t * KT)
        slot = None if (k_lo == 0 and k_hi == KT) else (0 if k_lo != 0 else 1)
        out.append((t, k_lo, k_hi, slot))
        r = t * KT + k_hi
    return out


def contributors(T, KT, P, t):
    """The kernel's own computation of [c_first, c_last] for tile t."""
    tb, te = t * KT, (t + 1) * KT
    c_first = (tb * P) // (T * KT)
    while sk_begin(T, KT, P, c_first + 1) <= tb:
        c_first += 1
    while c_first > 0 and sk_begin(T, KT, P, c_first) > tb:
        c_first -= 1
    c_last = c_first
    while c_last + 1 < P and sk_begin(T, KT, P, c_last + 1) < te:
        c_last += 1
    return c_first, c_last


@pytest.mark.parametrize("T,KT,P", [(128, 256, 148), (64, 256, 148), (32, 256, 148), (16, 64, 148), (32, 128, 148),
                                     (592, 256, 148), (5, 2048, 148), (33, 96, 148), (64, 256, 132), (7, 37, 148)])
def test_schedule_covers_every_k_step_once_and_fixup_reads_the_right_slots(T, KT, P):
    _check_schedule(T, KT, P)


def test_schedule_properties_on_random_shapes():
    """The same properties for random tile counts, k-step counts and CTA counts (hypothesis), including the
    degenerate corners: one tile, one k-step per tile, more tiles than CTAs, a single CTA."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(1, 300), st.integers(1, 260), st.integers(1, 160))
    def run(T, KT, P):
        if T * KT < P:
            P = T * KT
        _check_schedule(T, KT, P)

    run()


def _check_schedule(T, KT, P):
    assert T * KT >= P          # the host only picks the schedule then
    cover = {}
    slot_of = {}
    for c in range(P):
        segs = segments(T, KT, P, c)
        used = [s for (_, _, _, s) in segs if s is not None]
        assert len(used) == len(set(used)) <= 2, "a CTA owns two workspace slots, each used at most once"
        for (t, k_lo, k_hi, slot) in segs:
            assert 0 <= t < T and 0 <= k_lo < k_hi <= KT
            for k in range(k_lo, k_hi):
                assert (t, k) not in cover
                cover[(t, k)] = c
            if slot is not None:
                slot_of[(t, c)] = slot
    assert len(cover) == T * KT, "every k-step of every tile exactly once"
    for t in range(T):
        owners = sorted({cover[(t, k)] for k in range(KT)})
        c_first, c_last = contributors(T, KT, P, t)
        assert owners == list(range(c_first, c_last + 1)), "the kernel's contributor range is the true one"
        if len(owners) == 1:
            assert (t, owners[0]) not in slot_of      # whole tile in registers: no fixup, no ticket
            continue
        tb = t * KT
        for cc in owners:
            # the rule the finishing CTA uses to find cc's partial
            slot_cc = 0 if sk_begin(T, KT, P, cc) > tb else 1
            assert slot_of[(t, cc)] == slot_cc
    # every partial tile expects exactly as many tickets as it has contributors
    for t in range(T):
        c_first, c_last = contributors(T, KT, P, t)
        n_partial = sum(1 for cc in range(c_first, c_last + 1) if (t, cc) in slot_of)
        assert n_partial in (0, c_last - c_first + 1)
