"""FrameEnv ingest (recnn_b200/data/env.py: ratings CSV + embedding pickle -> EnvBase) against the
reference's own ingest (recnn.data.dataset_functions.prepare_dataset + utils.make_items_tensor behind
Env.process_env, recnn/data/env.py:133-176) on a synthetic ML-20M-shaped table; fixture
tests/golden/ingest.npz produced by oracle/make_golden.py:run_ingest_case from the unmodified reference.
CPU only: the DataLoader side uses embed_batch=batch_frames, which touches no CUDA memory."""
import pickle

import numpy as np
import pandas as pd
import torch

import recnn_b200
from oracle import recnn_oracle as O
from tests._golden import load_golden


def _write_inputs(g, tmp_path):
    df = pd.DataFrame({c: g["csv." + c] for c in ("userId", "movieId", "rating", "timestamp")})
    df.to_csv(tmp_path / "ratings.csv", index=False)
    emb = {int(k): torch.from_numpy(v) for k, v in zip(g["keys"], g["emb"])}
    with open(tmp_path / "emb.pkl", "wb") as fh:
        pickle.dump(emb, fh)
    return recnn_b200.data.DataPath(str(tmp_path) + "/", "ratings.csv", "emb.pkl", use_cache=False)


def test_ingest_matches_the_reference(tmp_path):
    g = load_golden("ingest.npz")
    frame = int(g["frame_size"])
    env = recnn_b200.data.FrameEnv(_write_inputs(g, tmp_path), frame_size=frame, batch_size=4, num_workers=0,
                                   embed_batch=recnn_b200.data.batch_frames)
    # item table and id maps (utils.py:203-214)
    assert np.array_equal(env.base.embeddings.numpy(), g["table"])
    assert [env.base.id_to_key[i] for i in range(len(g["keys"]))] == g["keys"].tolist()
    assert all(env.base.key_to_id[k] == i for i, k in enumerate(g["keys"].tolist()))
    # which users survive (more than frame_size interactions), every user's time-ordered arrays
    train, test = env.base.train_user_dataset, env.base.test_user_dataset
    kept = set(train.users) | set(test.users)
    want = set(g["users"].tolist())
    assert kept <= want and len(want - kept) == 2           # the reference drops the 2 longest train users ([2:])
    lengths = {u: len(g["u%d.items" % u]) for u in want}
    assert all(lengths[u] > frame for u in want)
    longest_two = sorted(lengths.values())[-2:]
    assert all(lengths[u] >= min(longest_two) for u in want - kept) or len(test.users) > 0
    for ds in (train, test):
        ls = [lengths[u] for u in ds.users]
        assert ls == sorted(ls, reverse=True)               # sort_users_itemwise: longest first
        for i, u in enumerate(ds.users):
            rec = ds[i]
            assert np.array_equal(rec["items"], g["u%d.items" % u]) and rec["items"].dtype == np.int64
            assert np.array_equal(rec["rates"], g["u%d.ratings" % u])        # 2 * (r - 2.5), float64
            assert rec["sizes"] == lengths[u] and rec["users"] == u
    # users at or below frame_size never reach a dataset
    assert all(len(g["u%d.items" % u]) <= frame for u in set(g["all_users"].tolist()) - want)


def test_dataloader_batches_equal_the_oracle_collate(tmp_path):
    """FrameEnv.train_dataloader with embed_batch=batch_frames: every minibatch equals the oracle's collate of
    the same users (sliding windows, sizes, users); HistoryCSR built from the same dataset plans the same rows."""
    g = load_golden("ingest.npz")
    frame = int(g["frame_size"])
    env = recnn_b200.data.FrameEnv(_write_inputs(g, tmp_path), frame_size=frame, batch_size=5, num_workers=0,
                                   embed_batch=recnn_b200.data.batch_frames)
    ds = env.base.train_user_dataset
    csr = recnn_b200.data.HistoryCSR.from_dataset(ds, frame)
    pos_of = {u: i for i, u in enumerate(ds.users)}
    seen = 0
    for batch in env.train_dataloader:
        users = batch["users"].tolist()
        col = O.collate_users([ds[pos_of[u]] for u in users], frame)
        assert np.array_equal(batch["items"].numpy(), col["items"])
        assert np.array_equal(batch["ratings"].numpy().view(np.uint32), col["ratings"].view(np.uint32))
        assert np.array_equal(batch["sizes"].numpy(), col["sizes"])
        row_offsets, n_rows = csr.plan_users([pos_of[u] for u in users])
        assert n_rows == col["items"].shape[0] and np.array_equal(np.diff(row_offsets), col["sizes"] - frame)
        seen += len(users)
    assert seen == len(ds)
