"""FrameEnv and its helpers with the reference's constructor surface
(recnn/data/env.py:23-256): ``UserDataset``, ``EnvBase``, ``DataPath``, ``Env``, ``FrameEnv``.

Only the seam that feeds the hot path matters here -- the DataLoaders whose ``collate_fn`` turns user
histories into minibatches through the pluggable ``embed_batch`` -- so the offline ingest (ratings CSV ->
per-user arrays) is a plain pandas/numpy pass with the reference's semantics and none of its options
(modin, progress bars).  ``FrameEnv.from_user_dict`` builds an environment from in-memory data.

Recommended with this package: ``embed_batch=recnn_b200.data.batch_frames`` and ``num_workers>=0``; the
collate then returns only ids/ratings and the gather happens on the GPU inside the update step.  The
default (``batch_tensor_embeddings``) needs ``num_workers=0`` when the table lives on the GPU, because a
DataLoader worker process cannot launch kernels on the parent's CUDA context.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from . import utils


class UserDataset(Dataset):
    """user -> {items, rates, sizes, users}; ``users`` is the list of user ids in this split."""

    def __init__(self, users, user_dict):
        self.users = users
        self.user_dict = user_dict

    def __len__(self):
        return len(self.users)

    def __getitem__(self, idx):
        uid = self.users[idx]
        rec = self.user_dict[uid]
        items = rec["items"][:]
        return {"items": items, "rates": rec["ratings"][:], "sizes": items.shape[0], "users": uid}


class EnvBase:
    """Serialisable state of an environment (same five fields as the reference's EnvBase)."""

    def __init__(self):
        self.train_user_dataset = None
        self.test_user_dataset = None
        self.embeddings = None
        self.key_to_id = None
        self.id_to_key = None


class DataPath:
    def __init__(self, base: str, ratings: str, embeddings: str, cache: str = "", use_cache: bool = True):
        self.ratings = base + ratings
        self.embeddings = base + embeddings
        self.cache = base + cache
        self.use_cache = use_cache


def _user_histories(ratings_csv, key_to_id, frame_size):
    """ratings CSV (userId, movieId, rating, timestamp) -> users with more than ``frame_size``
    interactions and their time-ordered item-row / rating arrays; ratings mapped r -> 2*(r-2.5)."""
    import pandas as pd
    df = pd.read_csv(ratings_csv)
    df["rating"] = 2.0 * (df["rating"] - 2.5)
    df["movieId"] = df["movieId"].map(key_to_id)
    df = df.dropna(subset=["movieId"])
    df["movieId"] = df["movieId"].astype(np.int64)
    counts = df.groupby("userId").size()
    users = counts[counts > frame_size].sort_values(ascending=False).index
    user_dict = {}
    for uid, grp in df.sort_values("timestamp").groupby("userId"):
        user_dict[uid] = {"items": grp["movieId"].to_numpy(), "ratings": grp["rating"].to_numpy()}
    return list(users), user_dict


class Env:
    def __init__(self, path: DataPath, prepare_dataset=None, embed_batch=utils.batch_tensor_embeddings, **kwargs):
        self.base = EnvBase()
        self.embed_batch = embed_batch
        self.prepare_dataset = prepare_dataset
        self._kwargs = kwargs
        if path is None:
            return
        if path.use_cache and os.path.isfile(path.cache):
            self.load_env(path.cache)
        else:
            self.process_env(path)
            if path.use_cache:
                self.save_env(path.cache)

    def process_env(self, path: DataPath, frame_size=10, test_size=0.05, seed=None):
        with open(path.embeddings, "rb") as fh:
            emb = pickle.load(fh)
        self.base.embeddings, self.base.key_to_id, self.base.id_to_key = utils.make_items_tensor(emb)
        users, user_dict = _user_histories(path.ratings, self.base.key_to_id, frame_size)
        self._split(users, user_dict, test_size, seed)

    def _split(self, users, user_dict, test_size, seed=None, drop_longest=2):
        rng = np.random.default_rng(seed)
        users = list(users)
        perm = rng.permutation(len(users))
        n_test = int(np.ceil(len(users) * test_size)) if len(users) > 1 else 0
        test = [users[i] for i in perm[:n_test]]
        train = [users[i] for i in perm[n_test:]]
        by_len = lambda us: sorted(us, key=lambda u: -user_dict[u]["items"].shape[0])
        self.base.train_user_dataset = UserDataset(by_len(train)[drop_longest:], user_dict)   # reference: [2:]
        self.base.test_user_dataset = UserDataset(by_len(test), user_dict)

    def load_env(self, where: str):
        with open(where, "rb") as fh:
            self.base = pickle.load(fh)

    def save_env(self, where: str):
        with open(where, "wb") as fh:
            pickle.dump(self.base, fh)


class FrameEnv(Env):
    """Static-length (frame) environment: every sample is ``frame_size`` items + the next one."""

    def __init__(self, path, frame_size=10, batch_size=25, num_workers=1, *args, **kwargs):
        kwargs["frame_size"] = frame_size
        super().__init__(path, *args, **kwargs)
        self.frame_size = frame_size
        self.batch_size = batch_size
        self.num_workers = num_workers
        if self.base.train_user_dataset is not None:
            self._make_loaders()

    @classmethod
    def from_user_dict(cls, embeddings, user_dict, frame_size=10, batch_size=25, num_workers=0, test_size=0.05,
                       embed_batch=utils.batch_tensor_embeddings, seed=0):
        """Environment from an item table fp32[n_items, D] and {user: {"items": int64[L], "ratings": float[L]}}
        (item ids already are table rows)."""
        env = cls(None, frame_size, batch_size, num_workers, embed_batch=embed_batch)
        env.base.embeddings = embeddings
        env.base.key_to_id = env.base.id_to_key = None
        users = [u for u, r in user_dict.items() if r["items"].shape[0] > frame_size]
        env._split(users, user_dict, test_size, seed, drop_longest=0)
        env._make_loaders()
        return env

    def _make_loaders(self):
        mk = lambda ds: DataLoader(ds, batch_size=self.batch_size, shuffle=True, num_workers=self.num_workers,
                                   collate_fn=self.prepare_batch_wrapper)
        self.train_dataloader = mk(self.base.train_user_dataset)
        self.test_dataloader = mk(self.base.test_user_dataset) if len(self.base.test_user_dataset) else None

    def prepare_batch_wrapper(self, x):
        return utils.prepare_batch_static_size(x, self.base.embeddings, embed_batch=self.embed_batch,
                                               frame_size=self.frame_size)

    def train_batch(self):
        return next(iter(self.train_dataloader))

    def device_feed(self, device="cuda", split="train"):
        """The same users and table as the DataLoaders, resident on ``device``: a
        ``recnn_b200.data.DeviceFrameFeed`` whose ``epoch(batch_size)`` / ``batch(...)`` yield the
        minibatches ``train_dataloader`` would (frame form; windows cut by a kernel instead of a
        DataLoader worker), and whose ``sample(n_rows)`` draws constant-size minibatches."""
        from .feed import HistoryCSR, DeviceFrameFeed
        ds = self.base.train_user_dataset if split == "train" else self.base.test_user_dataset
        return DeviceFrameFeed(HistoryCSR.from_dataset(ds, self.frame_size), self.base.embeddings, device)

    def test_batch(self):
        return next(iter(self.test_dataloader))
