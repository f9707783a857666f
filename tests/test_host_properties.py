"""Property tests (hypothesis) of the host-side batching / sharding / prosody helpers -- CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from dissc_b200 import dist as ddist
from dissc_b200 import infer as pinf
from dissc_b200 import inference as inf

lengths_st = st.lists(st.integers(min_value=1, max_value=900), min_size=1, max_size=60)


@settings(max_examples=60, deadline=None)
@given(lengths_st, st.integers(min_value=1, max_value=8))
def test_shard_by_length_partitions_and_balances(lengths, world):
    shards = ddist.shard_by_length(lengths, world)
    assert len(shards) == world
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(len(lengths)))                       # every utterance exactly once
    sizes = [len(s) for s in shards]
    assert max(sizes) - min(sizes) <= 1                            # dealt round-robin
    sums = [sum(lengths[i] for i in s) for s in shards]
    assert max(sums) - min(sums) <= max(lengths)                   # serpentine deal: imbalance below one utterance


@settings(max_examples=60, deadline=None)
@given(lengths_st, st.integers(min_value=1, max_value=16), st.integers(min_value=900, max_value=20000))
def test_batches_by_length_partition_and_limits(lengths, max_batch, max_frames):
    batches = inf.batches_by_length(lengths, max_batch, max_frames)
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(len(lengths)))
    for b in batches:
        assert 1 <= len(b) <= max_batch
        L = lengths[b[0]]
        assert all(lengths[i] <= L for i in b)                     # first item is the longest: T of the padded batch
        assert len(b) == 1 or len(b) * L <= max_frames


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 99), st.integers(1, 6), st.integers(1, 9)), min_size=1, max_size=30))
def test_morph_seq_len_length_and_values(runs):
    # consecutive runs must differ in their unit, else groupby merges them
    units, pitch, lens = [], [], []
    prev = None
    for tok, n, target in runs:
        if tok == prev:
            tok = (tok + 1) % 100
        prev = tok
        units += [tok] * n
        pitch += list(np.linspace(100.0, 200.0, n) + tok)
        lens.append(target)
    out = pinf.morph_seq_len(units, pitch, lens)
    assert len(out) == sum(lens)
    assert set(np.round(out, 6)) <= set(np.round(pitch, 6)) | {0.0}   # nearest-neighbour resampling invents no values


@settings(max_examples=40, deadline=None)
@given(st.lists(st.floats(min_value=60.0, max_value=400.0), min_size=3, max_size=50), st.floats(100.0, 300.0),
       st.floats(5.0, 60.0), st.integers(0, 3))
def test_rescale_f0_moves_voiced_statistics_only(voiced, mean, std, n_unvoiced):
    voiced = np.asarray(voiced, dtype=np.float32)
    if np.std(voiced) < 1e-3:
        voiced = voiced + np.arange(len(voiced), dtype=np.float32)
    f0 = np.concatenate([np.zeros(n_unvoiced, np.float32), voiced, np.zeros(n_unvoiced, np.float32)])
    out = inf.rescale_f0(f0, mean, std)
    assert np.all(out[f0 == 0] == 0)
    v = out[f0 != 0].astype(np.float64)
    assert abs(v.mean() - mean) < 1e-2 * max(1.0, abs(mean))
    assert abs(v.std(ddof=1) - std) < 2e-2 * std
