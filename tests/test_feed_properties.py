"""Property tests (hypothesis) of the feed's host-side planning against the numpy oracle: random ragged user
histories, random minibatches, random window ids.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import recnn_oracle as O
from recnn_b200.data.feed import HistoryCSR


def make_users(lengths, seed):
    rng = np.random.default_rng(seed)
    return [{"items": rng.integers(0, 1000, size=n, dtype=np.int64), "rates": rng.standard_normal(n) * 3,
             "sizes": n, "users": 100 + 3 * i} for i, n in enumerate(lengths)]


lengths_st = st.lists(st.integers(min_value=0, max_value=40), min_size=1, max_size=30)


@settings(max_examples=60, deadline=None)
@given(extra=lengths_st, frame=st.integers(1, 12), seed=st.integers(0, 10 ** 6), data=st.data())
def test_plans_agree_with_the_oracle_collate(extra, frame, seed, data):
    lengths = [frame + 1 + e for e in extra]                 # FrameEnv only keeps users with more than `frame` items
    users = make_users(lengths, seed)
    csr = HistoryCSR([u["users"] for u in users], [u["items"] for u in users], [u["rates"] for u in users], frame)
    assert csr.n_windows == sum(n - frame for n in lengths)
    # a random minibatch (users may repeat, as nothing forbids it)
    pos = data.draw(st.lists(st.integers(0, len(users) - 1), min_size=1, max_size=12))
    row_offsets, n_rows = csr.plan_users(pos)
    col = O.collate_users([users[i] for i in pos], frame)
    assert n_rows == col["items"].shape[0] and row_offsets[0] == 0
    assert np.array_equal(np.diff(row_offsets), col["sizes"] - frame)
    # the CSR slices reproduce every window of the collate
    for b, i in enumerate(pos):
        lo = csr.offsets[i]
        for w in range(int(row_offsets[b + 1] - row_offsets[b])):
            r = row_offsets[b] + w
            assert np.array_equal(csr.items[lo + w:lo + w + frame + 1], col["items"][r])
            assert np.array_equal(csr.ratings[lo + w:lo + w + frame + 1].view(np.uint32), col["ratings"][r].view(np.uint32))
    # window ids: owner lookup through win_offsets == the oracle's row selection
    ids = np.asarray(data.draw(st.lists(st.integers(0, csr.n_windows - 1), min_size=1, max_size=20)), dtype=np.int64)
    rows = O.collate_rows(users, frame, ids)
    owner = np.searchsorted(csr.win_offsets, ids, side="right") - 1
    local = ids - csr.win_offsets[owner]
    assert np.array_equal(csr.user_ids[owner], rows["users"])
    assert np.array_equal(np.stack([csr.items[csr.offsets[o] + l:csr.offsets[o] + l + frame + 1] for o, l in zip(owner, local)]),
                          rows["items"])
    assert np.array_equal((local == csr.win_counts[owner] - 1).astype(np.float32), rows["done"])


@settings(max_examples=40, deadline=None)
@given(extra=lengths_st, batch_size=st.integers(1, 9), drop_last=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_epoch_plan_partitions_the_permutation(extra, batch_size, drop_last, seed):
    frame = 5
    lengths = [frame + 1 + e for e in extra]
    users = make_users(lengths, seed)
    csr = HistoryCSR([u["users"] for u in users], [u["items"] for u in users], [u["rates"] for u in users], frame)
    perm = np.random.default_rng(seed).permutation(len(users))
    starts, counts, flat, row_starts, n_rows = csr.plan_epoch(perm, batch_size, drop_last)
    covered = np.concatenate([perm[s:s + c] for s, c in zip(starts, counts)]) if len(starts) else np.zeros(0, dtype=np.int64)
    expect = len(users) - (len(users) % batch_size if drop_last else 0)
    assert covered.size == expect and np.array_equal(covered, perm[:expect])
    for s, c, rs, nr in zip(starts, counts, row_starts, n_rows):
        plan = flat[rs:rs + c + 1]
        assert plan[0] == 0 and plan[-1] == nr == sum(lengths[i] - frame for i in perm[s:s + c])
        assert np.all(np.diff(plan) >= 1)
