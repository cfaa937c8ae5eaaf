"""CPU test of the RecBole-style metric definitions in datamining_recblr_b200/evaluation.py against the oracle."""
import numpy as np
import torch

from oracle import bdlru_oracle as O


def test_metrics_from_record_matches_oracle_definitions():
    from datamining_recblr_b200.evaluation import metrics_from_record
    rng = np.random.default_rng(0)
    ids = np.stack([rng.permutation(100)[:20] for _ in range(500)])
    pos = rng.integers(0, 100, 500)
    rec = torch.tensor(np.concatenate([(ids == pos[:, None]).astype(np.int32), np.ones((500, 1), np.int32)], 1))
    got = metrics_from_record(rec, (10, 20), None)
    ref = O.eval_metrics(ids, pos)
    for k, v in ref.items():
        assert abs(got[k] - v) < 1e-12, k
    rounded = metrics_from_record(rec, (10, 20))
    assert rounded["hit@10"] == round(ref["hit@10"], 4)
    assert rounded["recall@20"] == rounded["hit@20"] and abs(rounded["precision@10"] - rounded["hit@10"] / 10) < 1e-4


class _DenseStubModel(torch.nn.Module):
    """Stands in for the model on CPU: mean-pooled history embedding as the user vector, dense scores, stable top-k."""

    def __init__(self, n_items, D, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = torch.nn.Parameter(torch.randn(n_items, D, generator=g, dtype=torch.float64))

    def full_sort_predict(self, inter):
        h = self.emb[inter["item_id_list"]]
        m = (inter["item_id_list"] > 0).unsqueeze(-1)
        return ((h * m).sum(1) / inter["item_length"].view(-1, 1)) @ self.emb.T

    def full_sort_topk(self, inter, k, mask_padding_item=True):
        s = self.full_sort_predict(inter).detach().numpy()
        v, i = O.topk_lowest_index(s, k, mask_col0=mask_padding_item)
        return torch.tensor(v), torch.tensor(i)


def test_evaluate_unseen_users_matches_reference_per_user_loop():
    """The reference procedure (run_with_unseen.py:196-261) restated literally: B = 1 full_sort_predict per user, drop the
    pad column, sklearn ndcg_score(k=10) + argpartition Hit@10 over the dense matrix, rows without a known target dropped."""
    from sklearn.metrics import ndcg_score
    from datamining_recblr_b200.evaluation import evaluate_unseen_users
    rng = np.random.default_rng(5)
    n_users, n_items, L = 150, 60, 7
    model = _DenseStubModel(n_items, 8, 1)
    lens = rng.integers(1, L + 1, n_users)
    lists = np.zeros((n_users, L), np.int64)
    for u in range(n_users):
        lists[u, :lens[u]] = rng.integers(1, n_items, lens[u])
    true = rng.integers(1, n_items, n_users)
    true[::9] = 0     # target unknown to the vocabulary: the reference skips these rows

    y_scores = np.zeros((n_users, n_items - 1))
    y_true = np.zeros((n_users, n_items - 1))
    with torch.no_grad():
        for u in range(n_users):
            inter = {"item_id_list": torch.tensor(lists[u:u + 1]), "item_length": torch.tensor(lens[u:u + 1])}
            y_scores[u] = model.full_sort_predict(inter)[0].numpy()[1:]
            if true[u] > 0:
                y_true[u, true[u] - 1] = 1
    valid = y_true.sum(axis=1) > 0
    ref_ndcg = ndcg_score(y_true[valid], y_scores[valid], k=10)
    hits = []
    for row_t, row_s in zip(y_true[valid], y_scores[valid]):
        top = np.argpartition(row_s, -10)[-10:]
        hits.append(1 if len(np.intersect1d(top, np.where(row_t == 1)[0])) > 0 else 0)
    got = evaluate_unseen_users(model, lists, lens, true, k=10, batch_size=32)
    assert abs(got["hit@10"] - np.mean(hits)) < 1e-12
    assert abs(got["ndcg@10"] - ref_ndcg) < 1e-12
    assert evaluate_unseen_users(model, lists, lens, np.zeros(n_users, np.int64)) == {"hit@10": 0.0, "ndcg@10": 0.0}
