"""Nearest-item retrieval over the embedding table (recnn/data/db_con.py:45-56,
examples/streamlit_demo.py:189-215).

The reference hands the actor's generated action to an external vector index (a
Milvus server, or faiss IndexFlatL2 / IndexFlatIP in the demo) and gets back
``topk`` item ids and distances.  Here the index IS the embedding table that is
already resident in HBM for the update step: ``search`` is one exact
``[n, dim] x [dim, n_items]`` contraction on the tensor cores plus an exact
top-k, both in librecnn_b200.so (recnn_retrieve_topk).  No server, no CPU path.

``ItemIndex(table, metric).search(vecs, topk)`` returns a ``SearchResult`` with
the reference's accessors ``.id(device)`` / ``.dist(device)``.
``MilvusConnection(env, ...)`` keeps the reference's constructor shape for code
written against recnn.data.db_con.
"""
from __future__ import annotations

import torch

from .. import _lib

_METRICS = {"L2": _lib.METRIC_L2, "IP": _lib.METRIC_IP, "COS": _lib.METRIC_COS}


class SearchResult:
    """recnn/data/db_con.py:5-13: ids int64 [n, topk] and distances fp32 [n, topk], best first."""

    def __init__(self, ids, dist):
        self.ids, self.distances = ids, dist

    def id(self, device=None):
        return self.ids if device is None else self.ids.to(device)

    def dist(self, device=None):
        return self.distances if device is None else self.distances.to(device)


class ItemIndex:
    """Exact top-k search over an item-embedding matrix resident on one GPU.

    metric: "L2" (squared Euclidean distance, smallest first -- faiss IndexFlatL2 / Milvus L2),
            "IP" (inner product, largest first), "COS" (cosine similarity, largest first).
    """

    def __init__(self, embeddings, metric="L2"):
        if metric not in _METRICS:
            raise ValueError("metric must be one of %s" % sorted(_METRICS))
        if not torch.is_tensor(embeddings):
            embeddings = torch.as_tensor(embeddings)
        if embeddings.device.type != "cuda":
            raise _lib.RecnnError("ItemIndex needs the embedding table on a CUDA device (got %s); there is no CPU path"
                                  % embeddings.device)
        self.table = embeddings.detach().to(torch.float32).contiguous()
        self.metric_name, self.metric = metric, _METRICS[metric]
        self.device = self.table.device
        self.n_items, self.dim = int(self.table.shape[0]), int(self.table.shape[1])
        self.norms = None
        if self.metric != _lib.METRIC_IP:
            self.norms = torch.empty(self.n_items, dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().recnn_item_norms(self.table.data_ptr(), self.n_items, self.dim, self.metric,
                                                       self.norms.data_ptr(), _lib.stream_ptr(self.device)))
        self._ws = None

    def search(self, search_vecs, topk=10, search_param=None):
        """search_vecs: [n, dim] (tensor / array on any device).  Returns SearchResult on the index's device."""
        q = torch.as_tensor(search_vecs)
        if q.dim() == 1:
            q = q.unsqueeze(0)
        if q.dim() != 2 or q.shape[1] != self.dim:
            raise ValueError("search_vecs must be [n, %d]" % self.dim)
        topk = int(topk)
        if not 1 <= topk <= min(64, self.n_items):
            raise ValueError("topk must be in [1, min(64, n_items)]")
        q = q.detach().to(device=self.device, dtype=torch.float32).contiguous()
        n = int(q.shape[0])
        ids = torch.empty(n, topk, dtype=torch.int64, device=self.device)
        dist = torch.empty(n, topk, dtype=torch.float32, device=self.device)
        if n == 0:
            return SearchResult(ids, dist)
        L = _lib.lib()
        need = L.recnn_retrieve_workspace_bytes(n, self.n_items, topk)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(L.recnn_retrieve_topk(q.data_ptr(), n, self.dim, self.table.data_ptr(), self.n_items,
                                             _lib.ptr(self.norms), self.metric, topk, ids.data_ptr(), dist.data_ptr(),
                                             self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr(self.device)))
        return SearchResult(ids, dist)


class MilvusConnection(ItemIndex):
    """Constructor shape of recnn.data.db_con.MilvusConnection(env, name=..., port=..., param=...): the collection
    is ``env.base.embeddings``; ``param["metric_type"]`` may be "L2" (default) or "IP".  No Milvus server is involved."""

    def __init__(self, env, name="movies_L2", port="19530", param=None):
        metric = "L2"
        if param and "metric_type" in param:
            mt = str(param["metric_type"]).upper()
            metric = "IP" if mt.endswith("IP") else "L2"
        emb = env.base.embeddings if hasattr(env, "base") else env
        super().__init__(emb, metric)
        self.name = name
        self.statuses = {}

    def get_log(self):
        return self.statuses
