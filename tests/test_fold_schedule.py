"""Host logic of the time-folded ResBlock-pair kernel (csrc/conv_pair_fold.cu), replayed on the CPU.

hg_fold_info (include/hifigan_b200.h) returns the tiling and the MMA schedule the kernel walks.  This test
executes that description in numpy exactly the way the kernel does — de-interleaved phase slabs with zero
fill outside the tensor and a zeroed tail inside the last block group, one matrix product per MMA group
into a [128, 128] accumulator whose column blocks hold the phases in reverse order, E1 writing xt phase
slabs, E2 storing whole folded rows of the kept range — and compares with a direct evaluation of
    y = c2(lrelu(c1(a) + b1)) + b2 + x              (reference hifi/models.py:90-94)
in float64.  No GPU needed: it pins the geometry (tile origins, halos, kept rows), the operand offsets and
the weight-block release order for every (C, k, d1) the generator uses.
"""
import ctypes

import numpy as np
import pytest

from tts_king_b200 import _native


def fold_info(C, k, d1, L):
    info = _native.HgFoldInfo()
    _native.check(_native.lib().hg_fold_info(C, k, d1, L, ctypes.byref(info)))
    return info


def direct_pair(a, x, w1, b1, w2, b2, d1, slope=0.1):
    L, C = a.shape
    k = w1.shape[0]
    ch = (k - 1) // 2

    def conv(inp, w, b, d):
        out = np.tile(b, (L, 1)).astype(np.float64)
        for j in range(k):
            off = (j - ch) * d
            lo, hi = max(0, -off), min(L, L - off)
            if hi > lo:
                out[lo:hi] += inp[lo + off:hi + off] @ w[j].T
        return out

    xt = conv(a, w1, b1, d1)
    xt = np.where(xt > 0, xt, xt * slope)
    return conv(xt, w2, b2, 1) + x


def replay(info, a, x, w1, b1, w2, b2, C, k, d1, slope=0.1):
    """Execute the kernel's dataflow from the host's description.  a, x: [L, C]; w: [k, C_out, C_in]."""
    L = a.shape[0]
    F, fdiv, rowb = info.f, info.fdiv, 2 * C
    assert F == 128 // C and fdiv == F * d1 and info.r_out % fdiv == 0 and info.delta % fdiv == 0
    slab_rows, xt_rows = info.slab_phase_bytes // rowb, info.xt_phase_bytes // rowb
    assert info.nb_slab * d1 <= slab_rows
    nblk_item = -(-L // fdiv)
    y = np.full((L, C), np.nan)
    held_trace = []
    for tile in range(-(-L // info.r_out)):
        g0 = tile * info.r_out - info.delta
        assert g0 % fdiv == 0
        blk_first = g0 // fdiv + info.blk_off
        # ---- slab: phase h, row blk*d1 + r  <-  time row (blk_first + blk)*fdiv + h*d1 + r; block groups outside
        # [0, nblk_item) are TMA zero fill, rows >= L inside the last one are zeroed by the producer warp
        slab = np.full((F, slab_rows, C), np.nan)
        for h in range(F):
            for blk in range(info.nb_slab):
                for r in range(d1):
                    g = (blk_first + blk) * fdiv + h * d1 + r
                    inside_map = 0 <= blk_first + blk < nblk_item
                    slab[h, blk * d1 + r] = a[g] if (inside_map and g < L) else 0.0
        # ---- G1
        acc = np.full((128, 128), np.nan)
        held = set()

        def run(ops, n_ops, src, w):
            nonlocal acc
            last_tap = None
            for o in range(n_ops):
                ph, row, tap, ntap, dcol, rel = ops[o][:6]
                A = src[ph, row:row + 128]                                  # [128, C]
                Bm = np.concatenate([w[tap + q] for q in range(ntap)], 0)   # [ntap*C, C]: consecutive taps along N
                held.update(range(tap, tap + ntap))
                prod = A @ Bm.T
                if o == 0:
                    acc[:, :] = 0.0  # the zero-operand MMA that clears the accumulator
                    last_tap = np.full(F, -1)
                # ascending u: every phase must see its taps in the order 0, 1, ..., k-1 (position-independent
                # summation order = bit-identical chunking and the same arithmetic as the N = C kernel)
                for q in range(ntap):
                    h = F - 1 - (dcol // C + q)
                    assert tap + q == last_tap[h] + 1, (o, h, tap + q, last_tap[h])
                    last_tap[h] = tap + q
                acc[:, dcol:dcol + ntap * C] += prod
                if rel:
                    held.discard(min(held))
                held_trace.append(len(held))
            assert not held, "every weight block released exactly once per conv"
            assert (last_tap == k - 1).all()

        ops1 = [list(info.ops1[i]) for i in range(info.n_ops1)]
        ops2 = [list(info.ops2[i]) for i in range(info.n_ops2)]
        run(ops1, info.n_ops1, slab, w1)
        # ---- E1: M row i, column block q = phase F-1-q -> xt row tau = fdiv*(i // d1) + h*d1 + i % d1
        xt = np.full((F, xt_rows, C), np.nan)
        for i in range(128):
            for q in range(F):
                h = F - 1 - q
                tau = fdiv * (i // d1) + h * d1 + i % d1
                v = acc[i, q * C:(q + 1) * C] + b1
                v = np.where(v > 0, v, v * slope)
                if not (0 <= g0 + tau < L):
                    v = np.zeros(C)
                xt[tau % F, tau // F] = v
        # ---- G2 + E2: M row i = folded output row (tile*r_out)/F + i; kept while inside the tile and the sequence
        run(ops2, info.n_ops2, xt, w2)
        q0 = tile * info.r_out // F
        for i in range(info.r_out // F):
            for q in range(F):
                h = F - 1 - q
                t = (q0 + i) * F + h
                if 0 <= t < L:
                    assert np.isnan(y[t]).all(), "every output row written once"
                    y[t] = acc[i, q * C:(q + 1) * C] + b2 + x[t]
    return y, held_trace


CASES = [(64, 3, 1), (64, 3, 3), (64, 3, 5), (64, 7, 1), (64, 7, 3), (64, 7, 5), (64, 11, 1), (64, 11, 3), (64, 11, 5),
         (32, 7, 1), (32, 7, 3), (32, 7, 5), (32, 11, 1), (32, 11, 3), (32, 11, 5), (32, 5, 2), (64, 5, 7)]


CASES16 = [(16, 3, 1), (16, 3, 3), (16, 3, 5), (16, 7, 1), (16, 7, 3), (16, 7, 5), (16, 11, 1), (16, 11, 3), (16, 11, 5), (16, 5, 2)]


@pytest.mark.parametrize("C,k,d1", CASES + CASES16)
@pytest.mark.parametrize("L", [4, 236, 1000, 1204])
def test_fold_schedule_reproduces_the_pair(C, k, d1, L):
    L = -(-L // (128 // C)) * (128 // C)  # the folded kernel needs L % F == 0 (F = 8 at 16 channels)
    info = fold_info(C, k, d1, L)
    assert info.fusable == 1
    assert info.smem_bytes <= 227 * 1024
    assert info.stages >= (2 * k if info.weights_resident else info.f + 2)
    assert info.weights_resident or C == 64
    rng = np.random.default_rng(C * 1000 + k * 10 + d1 + L)
    a = rng.standard_normal((L, C))
    x = rng.standard_normal((L, C))
    w1 = rng.standard_normal((k, C, C)) / (C * k) ** 0.5
    w2 = rng.standard_normal((k, C, C)) / (C * k) ** 0.5
    b1, b2 = rng.standard_normal(C) * 0.1, rng.standard_normal(C) * 0.1
    y, held = replay(info, a, x, w1, b1, w2, b2, C, k, d1)
    ref = direct_pair(a, x, w1, b1, w2, b2, d1)
    assert not np.isnan(y).any()
    assert np.abs(y - ref).max() <= 1e-10
    # streamed weights: the blocks held at any time must fit the ring
    if not info.weights_resident:
        assert max(held) <= info.stages - 1, (max(held), info.stages)


def test_fold_falls_back_where_it_does_not_apply():
    assert fold_info(64, 11, 1, 1001).fusable == 0   # L not a multiple of F = 2
    assert fold_info(32, 3, 1, 1000).fusable == 0    # no kernel instantiated (HBM-bound on the N = C kernel anyway)
    assert fold_info(128, 3, 1, 1000).fusable == 0   # only C = 16 / 32 / 64
    assert fold_info(16, 7, 3, 1004).fusable == 0    # 1004 % 8 != 0
    assert fold_info(32, 11, 5, 1002).fusable == 0   # 1002 % 4 != 0


RING_CASES = [(5, 5), (7, 7), (7, 5), (7, 4), (9, 6), (9, 5), (9, 4), (11, 6), (11, 5), (11, 4)]


def _ring(k, s, tap, cpar):
    slot, par, mir, slots = (ctypes.c_int32() for _ in range(4))
    _native.check(_native.lib().hg_fold_ring_query(k, s, tap, cpar, ctypes.byref(slot), ctypes.byref(par), ctypes.byref(mir),
                                                  ctypes.byref(slots)))
    return slot.value, par.value, bool(mir.value), slots.value


@pytest.mark.parametrize("k,s", RING_CASES)
def test_streamed_weight_ring_plan(k, s):
    """The streamed-weight ring of the folded pair kernel has a COMPILE-TIME period (conv_pair_fold.cu::FoldRingPlan):
    slot, barrier and parity of every tap are constants of the unrolled code plus one run-time bit (the parity of the
    conv's position in the ring order).  hg_fold_ring_query evaluates the kernel's own constexpr functions; this test
    replays producer and consumer over many convs with the mbarrier phase rule (a wait with parity P passes once the
    phase of parity P has completed, phases complete strictly in order):
      * every use of a slot's `full` barrier waits for exactly the next phase (parities alternate 0, 1, 0, ... per
        slot), so a consumer can never pass on a stale phase, and the producer's `empty` wait (parity ^ 1) passes at
        once for the first fill and thereafter needs exactly the previous occupant's release;
      * the two taps of every MMA group's run are contiguous in shared memory — neighbouring slots, or slot S - 1
        followed by the mirror slot, which then holds the right tap;
      * the blocks fit the slots the host allocates for that period."""
    full_phase = [0] * s    # completed phases of w_full[slot]
    empty_phase = [0] * s   # completed phases of w_empty[slot]
    held = [None] * (s + 1)  # tap held by each physical slot (index s = the mirror)
    slots_expected = s + (1 if k > s else 0)
    for conv in range(9):  # G1(0) G1(1) G2(0) G1(2) ...: only the parity of the position enters
        cpar = conv & 1
        # producer: taps in order; consumer: group oi uses the run (oi - 1, oi) and then releases tap oi - 1
        for tap in range(k):
            slot, par, mirror, slots = _ring(k, s, tap, cpar)
            assert slots == slots_expected and slot == tap % s
            assert mirror == (k > s and tap >= s and tap % s == 0)
            # producer waits for the release of the slot's previous occupant: phase (uses so far - 1) of w_empty
            uses_so_far = full_phase[slot]
            assert (par ^ 1) == ((uses_so_far - 1) & 1) if uses_so_far else (par ^ 1) == 1
            assert empty_phase[slot] == uses_so_far, "the previous occupant must have been released before the refill"
            held[slot] = tap
            if mirror:
                held[s] = tap
            # consumer waits for this fill: it is phase `uses_so_far` of w_full, and the wait names exactly that parity
            assert par == (uses_so_far & 1), (tap, slot, par, uses_so_far)
            full_phase[slot] += 1
            # MMA group `tap` (oi = tap >= 1) multiplies the run (tap - 1, tap): contiguous blocks
            if tap >= 1:
                first = (tap - 1) % s
                second = first + 1  # physical slot right behind it (s = the mirror)
                assert held[first] == tap - 1 and held[second] == tap, (tap, held)
                empty_phase[first] += 1  # ... and releases tap - 1
        empty_phase[(k - 1) % s] += 1  # the last group (oi = k) uses tap k - 1 alone and releases it


def test_streamed_weight_ring_rejects_unknown_periods():
    for k, s in ((11, 7), (7, 6), (5, 4), (3, 3), (11, 0)):
        slot = ctypes.c_int32()
        assert _native.lib().hg_fold_ring_query(k, s, 0, 0, ctypes.byref(slot), ctypes.byref(slot), ctypes.byref(slot),
                                                ctypes.byref(slot)) != 0
