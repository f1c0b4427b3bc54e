"""Nearest-item retrieval: oracle pinned against the reference's own ranking definition
(examples/streamlit_demo.py:207-215 ranks items by a scipy.spatial.distance metric against the generated
action; :189-203 builds faiss IndexFlatL2 / IndexFlatIP / normalised-IP indexes), CUDA path against the oracle."""
import numpy as np
import pytest
import torch

from oracle import recnn_oracle as O


def _case(seed, n_items, dim, n_q):
    rng = np.random.default_rng(seed)
    table = rng.standard_normal((n_items, dim)).astype(np.float32)
    q = (table[rng.integers(0, n_items, n_q)] * 0.5 + 0.5 * rng.standard_normal((n_q, dim))).astype(np.float32)
    return table, q


def test_oracle_matches_scipy_ranking():
    """The reference's rank(): sorted(items, key=metric(item_embedding, gen_action))[:k]."""
    from scipy.spatial import distance
    table, q = _case(1, 400, 16, 5)
    ids_l2, d_l2 = O.retrieve_topk(q, table, 7, "L2")
    ids_cos, d_cos = O.retrieve_topk(q, table, 7, "COS")
    ids_ip, d_ip = O.retrieve_topk(q, table, 7, "IP")
    for i in range(q.shape[0]):
        eu = np.asarray([distance.euclidean(t, q[i]) for t in table.astype(np.float64)])
        want = np.argsort(eu, kind="stable")[:7]
        assert np.array_equal(ids_l2[i], want)
        np.testing.assert_allclose(d_l2[i], eu[want] ** 2, rtol=1e-6, atol=1e-6)      # faiss / Milvus L2 = squared
        co = np.asarray([distance.cosine(t, q[i]) for t in table.astype(np.float64)])
        want = np.argsort(co, kind="stable")[:7]
        assert np.array_equal(ids_cos[i], want)
        np.testing.assert_allclose(d_cos[i], 1.0 - co[want], rtol=1e-6, atol=1e-6)
        ip = table.astype(np.float64) @ q[i].astype(np.float64)
        assert np.array_equal(ids_ip[i], np.argsort(-ip, kind="stable")[:7])


def test_oracle_ties_go_to_the_smaller_id():
    table = np.zeros((6, 4), np.float32)
    table[[1, 4]] = 1.0
    ids, _ = O.retrieve_topk(np.ones((1, 4), np.float32), table, 3, "L2")
    assert ids.tolist() == [[1, 4, 0]]


def _check(ids, dist, table, q, k, metric):
    want_ids, want_d = O.retrieve_topk(q, table, k, metric)
    scale = max(1.0, float(np.abs(want_d).max()))
    np.testing.assert_allclose(dist, want_d, rtol=2e-5, atol=2e-5 * scale)
    # ids: identical except where two candidates are closer than the fp32 accuracy of the scores
    full_ids, full_d = O.retrieve_topk(q, table, min(table.shape[0], k + 8), metric)
    for i in range(q.shape[0]):
        if np.array_equal(ids[i], want_ids[i]):
            continue
        lookup = dict(zip(full_ids[i].tolist(), full_d[i].tolist()))
        for r in range(k):
            assert int(ids[i, r]) in lookup, (i, r, ids[i], want_ids[i])
            assert abs(lookup[int(ids[i, r])] - want_d[i, r]) <= 2e-5 * scale, (i, r)
        assert len(set(ids[i].tolist())) == k


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["L2", "IP", "COS"])
@pytest.mark.parametrize("n_items,dim,n_q,k", [
    (26744, 128, 1, 10),        # one query against the ML-20M-shaped table (serving)
    (26744, 128, 300, 10),      # a batch (column ranges not split)
    (1000, 128, 5, 64),         # largest k
    (50, 7, 4, 3),              # odd width: CUDA-core contraction
    (333, 32, 40, 16),
])
def test_search_vs_oracle(metric, n_items, dim, n_q, k):
    import recnn_b200
    table, q = _case(n_items + dim, n_items, dim, n_q)
    index = recnn_b200.data.ItemIndex(torch.from_numpy(table).cuda(), metric)
    res = index.search(torch.from_numpy(q), topk=k)
    ids, dist = res.id("cpu").numpy(), res.dist("cpu").numpy()
    assert ids.shape == (n_q, k) and ids.dtype == np.int64 and dist.dtype == np.float32
    _check(ids, dist, table, q, k, metric)


@pytest.mark.gpu
def test_search_exact_duplicates_and_self_match():
    import recnn_b200
    table, _ = _case(3, 500, 128, 1)
    table[123] = table[7]                                   # exact duplicate rows: tie -> smaller id first
    index = recnn_b200.data.ItemIndex(torch.from_numpy(table).cuda(), "L2")
    res = index.search(torch.from_numpy(table[[7, 200]]), topk=3)
    ids, dist = res.id("cpu").numpy(), res.dist("cpu").numpy()
    assert ids[0, 0] == 7 and ids[0, 1] == 123 and ids[1, 0] == 200
    assert dist[0, 0] <= 1e-3 and dist[1, 0] <= 1e-3 and np.all(dist >= 0)


@pytest.mark.gpu
def test_actor_output_feeds_the_index_like_the_demo():
    """streamlit_demo.py:355-366: action = actor(state); D, I = index.search(action, k)."""
    import recnn_b200
    torch.manual_seed(0)
    actor = recnn_b200.nn.Actor(1290, 128, 256).cuda().eval()
    table, _ = _case(9, 2000, 128, 1)
    state = torch.randn(17, 1290)
    action = actor(state)
    res = recnn_b200.data.db_con.MilvusConnection(type("E", (), {"base": type("B", (), {"embeddings": torch.from_numpy(table).cuda()})})()).search(action, topk=10)
    _check(res.id("cpu").numpy(), res.dist("cpu").numpy(), table, action.cpu().numpy(), 10, "L2")


def test_index_rejects_cpu_tables():
    import recnn_b200
    with pytest.raises(recnn_b200._lib.RecnnError):
        recnn_b200.data.ItemIndex(torch.zeros(4, 4), "L2")
