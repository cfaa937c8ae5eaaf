"""CPU tests of the data path (datamining_recblr_b200/data.py) against a plain-Python restatement of RecBole's
sequential augmentation + leave-one-out rule."""
import numpy as np
import torch

from datamining_recblr_b200 import data as D


def _python_reference(users, items, times, L):
    per_user = {}
    for u, i, t in sorted(zip(users, items, times), key=lambda x: (x[0], x[2])):
        per_user.setdefault(u, []).append(i)
    out = {"train": [], "valid": [], "test": []}
    for u, seq in per_user.items():
        n = len(seq)
        for i in range(1, n):
            hist = seq[max(0, i - L):i]
            row = (u, hist + [0] * (L - len(hist)), len(hist), seq[i])
            out["test" if i == n - 1 else ("valid" if i == n - 2 else "train")].append(row)
    return out


def test_build_sequences_matches_python_reference():
    rng = np.random.default_rng(0)
    n = 4000
    users = rng.integers(1, 60, n)
    items = rng.integers(1, 40, n).astype(np.int32)
    times = rng.random(n)
    L = 7
    train, valid, test, n_items = D.build_sequences(users, items, times, L)
    ref = _python_reference(users.tolist(), items.tolist(), times.tolist(), L)
    assert n_items == items.max() + 1
    for name, arr in (("train", train), ("valid", valid), ("test", test)):
        got = sorted((int(u), tuple(h.tolist()), int(l), int(t)) for u, h, l, t in zip(arr.user, arr.hist, arr.length, arr.target))
        exp = sorted((u, tuple(h), l, t) for u, h, l, t in ref[name])
        assert got == exp, name
    assert len(test) == len(np.unique(users[np.isin(users, [u for u in np.unique(users) if (users == u).sum() >= 2])]))


def test_k_core_and_tokenize():
    users = np.array([1] * 6 + [2] * 5 + [3] * 2 + [4] * 5)
    items = np.array([10, 11, 12, 13, 14, 15, 10, 11, 12, 13, 14, 99, 98, 10, 11, 12, 13, 14])
    keep = D.k_core_filter(users, items, 5, 3)
    assert set(users[keep]) == {1, 2, 4} and 99 not in items[keep] and 15 not in items[keep]
    ids, vocab = D.tokenize(items[keep])
    assert ids.min() == 1 and vocab[0] is None and len(vocab) == ids.max() + 1


def test_inter_file_round_trip_and_loader(tmp_path):
    path = str(tmp_path / "toy.inter")
    n = D.write_synthetic_inter(path, n_users=80, n_items=50, mean_len=12, seed=1)
    users, items, times = D.read_inter(path)
    assert len(users) == n
    train, valid, test, n_items, vocab = D.load_dataset(path, max_len=10)
    assert len(valid) > 0 and len(test) == len(valid) and n_items == len(vocab)
    assert train.hist.shape[1] == 10 and (train.hist[np.arange(len(train)), train.length - 1] > 0).all()
    assert ((train.hist > 0).sum(1) == train.length).all()              # right padding only
    loader = D.PinnedBatchLoader(train, batch_size=64, device="cpu", shuffle=True, seed=3)
    seen = 0
    for b in loader:
        assert set(b) == {"item_id_list", "item_length", "item_id"}
        assert b["item_id_list"].dtype == torch.int64 and b["item_id_list"].shape[1] == 10
        seen += b["item_id"].shape[0]
    assert seen == len(train) and len(loader) == -(-len(train) // 64)
    a = [b["item_id"].clone() for b in D.PinnedBatchLoader(train, 64, "cpu", shuffle=True, seed=3)]
    c = [b["item_id"].clone() for b in D.PinnedBatchLoader(train, 64, "cpu", shuffle=True, seed=3)]
    assert all(torch.equal(x, y) for x, y in zip(a, c))                  # deterministic for a fixed seed
