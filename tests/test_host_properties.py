"""Property tests (hypothesis) for the host-side planners: the utterance sharder, the time-chunk planner
and the length-bucket planner.  CPU only."""
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import fixtures as fx
from tts_king_b200 import parallel, ragged

lengths_st = st.lists(st.integers(min_value=1, max_value=5000), min_size=0, max_size=60)


@settings(max_examples=150, deadline=None)
@given(lengths=lengths_st, world=st.integers(min_value=1, max_value=9))
def test_shard_utterances_is_a_balanced_partition(lengths, world):
    shards = parallel.shard_utterances(lengths, world)
    assert len(shards) == world
    assert sorted(i for s in shards for i in s) == list(range(len(lengths)))       # every utterance exactly once
    assert shards == parallel.shard_utterances(lengths, world)                    # deterministic
    loads = [sum(lengths[i] for i in s) for s in shards]
    if lengths:
        # greedy longest-first: no rank exceeds the lightest one by more than one (the longest) utterance
        assert max(loads) - min(loads) <= max(lengths)
    assert all(s == sorted(s) for s in shards)


@settings(max_examples=150, deadline=None)
@given(T=st.integers(min_value=0, max_value=400000), parts=st.integers(min_value=1, max_value=16),
       halo=st.integers(min_value=0, max_value=40))
def test_plan_time_chunks_covers_exactly_once(T, parts, halo):
    chunks = parallel.plan_time_chunks(T, parts, halo)
    assert len(chunks) == parts
    assert sum(c.frames for c in chunks) == T
    sizes = [c.frames for c in chunks]
    assert max(sizes) - min(sizes) <= 1
    pos = 0
    for c in chunks:
        assert c.start == pos and c.stop >= c.start
        pos = c.stop
        if c.frames:
            assert c.lo == max(0, c.start - halo) and c.hi == min(T, c.stop + halo)   # halos are clipped at the true ends
            assert 0 <= c.lo <= c.start and c.stop <= c.hi <= T
    assert pos == T


@settings(max_examples=100, deadline=None)
@given(frames=st.lists(st.integers(min_value=1, max_value=2000), min_size=1, max_size=40),
       halo=st.integers(min_value=0, max_value=20), launch_cost=st.integers(min_value=0, max_value=3000))
def test_length_buckets_are_optimal_among_the_obvious_plans(frames, halo, launch_cost):
    t_max = max(frames)
    buckets = ragged.plan_length_buckets(frames, halo, t_max, launch_cost)
    assert sorted(i for b in buckets for i in b) == list(range(len(frames)))

    def cost(plan):
        return sum(len(b) * ragged.bucket_extent(frames, b, halo, t_max) + launch_cost for b in plan)

    c = cost(buckets)
    assert c <= cost([list(range(len(frames)))])                 # never worse than the padded batch
    assert c <= cost([[i] for i in range(len(frames))])          # nor than one forward per utterance
    # buckets are contiguous in length order and come longest first
    tops = [max(frames[i] for i in b) for b in buckets]
    lows = [min(frames[i] for i in b) for b in buckets]
    assert tops == sorted(tops, reverse=True)
    assert all(lows[k] >= tops[k + 1] for k in range(len(buckets) - 1))


@pytest.mark.parametrize("cfg,reach,halo", [(fx.V1, 3258, 13), (fx.V2_NARROW, 3258, 13), (fx.TINY_RB1, 3258, 13)])
def test_receptive_reach_of_the_known_configs(cfg, reach, halo):
    h = fx.make_h(cfg)
    assert parallel.receptive_reach_samples(h) == reach
    assert parallel.halo_frames(h) == halo
    assert parallel.hop_length(h) == 256


def test_receptive_reach_resblock2():
    h = fx.make_h(fx.V3_RB2)
    # ResBlock2: one conv per dilation, no dilation-1 second conv (hifi/models.py:104-143)
    r = parallel.receptive_reach_samples(h)
    assert r > 0 and parallel.halo_frames(h) == -(-r // parallel.hop_length(h))
