"""Data path feeding the hot path: what RecBole's `SequentialDataset` + `TrainDataLoader` /
`FullSortEvalDataLoader` do for this model ([upstream], restated from behaviour — Appendix D), as flat pre-tokenised arrays
in pinned host memory with double-buffered asynchronous H2D copies.  At B200 speeds an ML-1M epoch is < 1 s of GPU time, so
a pandas/CPU collate per batch (RecBole runs `worker = 0`) becomes the bottleneck; here a batch is one gather from a
pinned int32 matrix and one non-blocking copy issued a batch ahead on a side stream.

On-disk format: RecBole "atomic" interaction file, tab separated, header `user_id:token  item_id:token  timestamp:float`
(reference usage: trim.py:3-4, run_with_unseen.py:48-66).  Pipeline, mirroring the reference configs (config.yaml:26-27):
k-core filter -> token ids (0 = [PAD]) -> per-user time order -> data augmentation (sample i = previous <= L items,
RIGHT-padded with 0, + target item i) -> leave-one-out split (last sample -> test, second last -> valid, rest -> train).
"""
import numpy as np
import torch


# ----------------------------------------------------------------------------- file -> arrays
def read_inter(path, user_field="user_id", item_field="item_id", time_field="timestamp"):
    """Reads a RecBole atomic `.inter` file; returns (users, items, times) as numpy arrays of raw tokens / floats."""
    import pandas as pd
    df = pd.read_csv(path, sep="\t")
    cols = {c.split(":")[0]: c for c in df.columns}
    t = df[cols[time_field]].to_numpy(dtype=np.float64) if time_field in cols else np.arange(len(df), dtype=np.float64)
    return df[cols[user_field]].to_numpy(), df[cols[item_field]].to_numpy(), t


def k_core_filter(users, items, min_user_inter=5, min_item_inter=5):
    """Iteratively drops users / items with fewer than the given number of interactions (RecBole's
    `user_inter_num_interval: [5, inf)` / `item_inter_num_interval`).  Returns a boolean keep-mask."""
    keep = np.ones(len(users), dtype=bool)
    while True:
        _, uinv, ucnt = np.unique(users[keep], return_inverse=True, return_counts=True)
        _, iinv, icnt = np.unique(items[keep], return_inverse=True, return_counts=True)
        bad = (ucnt[uinv] < min_user_inter) | (icnt[iinv] < min_item_inter)
        if not bad.any():
            return keep
        idx = np.flatnonzero(keep)
        keep[idx[bad]] = False


def tokenize(raw):
    """Raw tokens -> contiguous ids starting at 1 (0 is [PAD]); returns (ids int32, vocabulary with vocab[0] = None)."""
    vocab, inv = np.unique(raw, return_inverse=True)
    return (inv + 1).astype(np.int32), np.concatenate([[None], vocab.astype(object)])


class SequenceArrays:
    """One split as flat arrays: hist int32 [n, L] (right-padded with 0), length int32 [n], target int32 [n], user int32 [n]."""

    def __init__(self, hist, length, target, user):
        self.hist, self.length, self.target, self.user = hist, length, target, user

    def __len__(self):
        return len(self.target)


def build_sequences(user_ids, item_ids, times, max_len):
    """Augmentation + leave-one-out split.  Returns (train, valid, test, n_items incl. PAD).
    For a user with time-ordered items s_0..s_{n-1} the samples are (hist = s_{max(0,i-L)}..s_{i-1}, target = s_i),
    i = 1..n-1; the last one is the test sample, the one before the validation sample, the rest train (users with too
    few interactions simply contribute fewer samples, like RecBole)."""
    order = np.lexsort((times, user_ids))          # by user, then time (stable for equal timestamps)
    u, it = user_ids[order], item_ids[order].astype(np.int32)
    starts = np.flatnonzero(np.r_[True, u[1:] != u[:-1]])
    counts = np.diff(np.r_[starts, len(u)])
    # sample (user k, position i) for i in 1..n_k-1, enumerated flat
    n_samp = np.maximum(counts - 1, 0)
    total = int(n_samp.sum())
    seg = np.repeat(np.arange(len(starts)), n_samp)                 # user index of each sample
    first = np.cumsum(n_samp) - n_samp
    i = np.arange(total) - first[seg] + 1                           # target position inside the user's sequence
    begin = np.maximum(i - max_len, 0)
    length = (i - begin).astype(np.int32)
    cols = np.arange(max_len)[None, :]
    src = starts[seg][:, None] + begin[:, None] + cols
    valid = cols < length[:, None]
    hist = np.where(valid, it[np.minimum(src, len(it) - 1)], 0).astype(np.int32)
    target = it[starts[seg] + i]
    user = u[starts[seg]].astype(np.int32)
    last = i == counts[seg] - 1
    second_last = i == counts[seg] - 2

    def take(mask):
        return SequenceArrays(hist[mask], length[mask], target[mask], user[mask])

    return take(~last & ~second_last), take(second_last), take(last), int(item_ids.max()) + 1


def load_dataset(path, max_len, min_user_inter=5, min_item_inter=5):
    """`.inter` file -> (train, valid, test, n_items, item vocabulary)."""
    users, items, times = read_inter(path)
    keep = k_core_filter(users, items, min_user_inter, min_item_inter)
    uid, _ = tokenize(users[keep])
    iid, vocab = tokenize(items[keep])
    train, valid, test, n_items = build_sequences(uid, iid, times[keep], max_len)
    return train, valid, test, n_items, vocab


# ----------------------------------------------------------------------------- arrays -> device batches
class PinnedBatchLoader:
    """Yields RecBole-style interactions `{item_id_list [B, L] int64, item_length [B] int64, item_id [B] int64}` on
    `device`.  The split lives in pinned host memory; batch b+1 is gathered and copied on a side stream while batch b
    is being consumed (the consumer's stream waits on the copy's event, no host sync)."""

    def __init__(self, arrays, batch_size, device, shuffle=False, seed=2020, drop_last=False,
                 item_field="item_id", list_suffix="_list", length_field="item_length"):
        self.a, self.batch_size, self.device = arrays, batch_size, torch.device(device)
        self.shuffle, self.drop_last = shuffle, drop_last
        self.rng = np.random.default_rng(seed)
        self.keys = (item_field + list_suffix, length_field, item_field)
        self.hist = torch.from_numpy(np.ascontiguousarray(arrays.hist))
        self.length = torch.from_numpy(np.ascontiguousarray(arrays.length))
        self.target = torch.from_numpy(np.ascontiguousarray(arrays.target))
        self.cuda = self.device.type == "cuda"
        if self.cuda:
            self.hist, self.length, self.target = self.hist.pin_memory(), self.length.pin_memory(), self.target.pin_memory()
            self.stream = torch.cuda.Stream(device=self.device)
            L = arrays.hist.shape[1]
            self.stage = [(torch.empty((batch_size, L), dtype=torch.int32).pin_memory(),
                           torch.empty(batch_size, dtype=torch.int32).pin_memory(),
                           torch.empty(batch_size, dtype=torch.int32).pin_memory()) for _ in range(2)]
            self.slot_done = [None, None]   # event of the last H2D copy issued FROM each staging slot

    def __len__(self):
        n = len(self.a)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def _index_batches(self):
        n = len(self.a)
        order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        for s in range(0, n, self.batch_size):
            idx = order[s:s + self.batch_size]
            if len(idx) < self.batch_size and self.drop_last:
                return
            yield torch.from_numpy(idx)

    def _issue(self, idx, slot):
        """Gather into the pinned staging buffers and start the H2D copy on the side stream."""
        n = len(idx)
        if not self.cuda:
            return (self.hist[idx].long(), self.length[idx].long(), self.target[idx].long()), None
        h, l, t = self.stage[slot]
        if self.slot_done[slot] is not None:
            self.slot_done[slot].synchronize()   # the copy that last read this pinned slot must be over before regathering
        torch.index_select(self.hist, 0, idx, out=h[:n])
        torch.index_select(self.length, 0, idx, out=l[:n])
        torch.index_select(self.target, 0, idx, out=t[:n])
        with torch.cuda.stream(self.stream):
            dev = tuple(x[:n].to(self.device, non_blocking=True).long() for x in (h, l, t))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.slot_done[slot] = ev
        return dev, ev

    def __iter__(self):
        pending, slot = None, 0
        for idx in self._index_batches():
            nxt = self._issue(idx, slot)
            slot ^= 1
            if pending is not None:
                yield self._finish(pending)
            pending = nxt
        if pending is not None:
            yield self._finish(pending)

    def _finish(self, pending):
        dev, ev = pending
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            for x in dev:
                x.record_stream(torch.cuda.current_stream(self.device))
        return dict(zip(self.keys, dev))


def write_synthetic_inter(path, n_users, n_items, mean_len, seed=0):
    """Writes a synthetic atomic `.inter` file with the statistics of a named dataset (the real files are absent:
    .MISSING_LARGE_BLOBS).  Sequence lengths are geometric-ish around mean_len (>= 5), items Zipf-distributed."""
    rng = np.random.default_rng(seed)
    lens = np.maximum(5, rng.geometric(1.0 / max(mean_len - 4, 1), size=n_users) + 4)
    users = np.repeat(np.arange(1, n_users + 1), lens)
    ranks = rng.zipf(1.2, size=len(users))
    items = (ranks - 1) % n_items + 1
    times = rng.random(len(users)) + np.repeat(np.arange(n_users), lens) * 0.0 + np.concatenate([np.arange(n) for n in lens])
    with open(path, "w") as fh:
        fh.write("user_id:token\titem_id:token\ttimestamp:float\n")
        for u, i, t in zip(users, items, times):
            fh.write(f"{u}\t{i}\t{t:.6f}\n")
    return len(users)
