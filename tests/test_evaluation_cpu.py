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
