"""Full-sort evaluation on top of the fused scorer — the RecBole eval boundary of SURVEY.md §8 f2.

RecBole's stock path ([upstream] `Trainer._full_sort_batch_eval` + `Collector.eval_batch_collect`) needs the
dense `[B, n_items]` score matrix: it writes `scores[:, 0] = -inf`, runs `torch.topk(scores, max(topk))`, builds an int
`pos_matrix` of the same size and gathers it at the top-k ids.  Here the model's `full_sort_topk` returns the `[B, k]` ids
directly (never forming `[B, n_items]`), and the collector's "rec.topk" record — `cat(pos_idx [B, K], pos_len [B, 1])` —
is rebuilt from the ids by comparison with the positive item, so RecBole's `Evaluator` (or the metric functions below,
which restate its Hit / NDCG / MRR / Recall / Precision definitions for one positive per user) can consume it unchanged.
"""
import torch


def topk_record(model, interaction, positive_i, k):
    """The "rec.topk" tensor of RecBole's Collector for one eval batch: int [B, k + 1] = (hit flags at ranks 1..k,
    number of positives).  positive_i [B] holds the single held-out item of each user (sequential leave-one-out eval:
    FullSortEvalDataLoader yields positive_u = arange(B), positive_i = interaction[item_id], no history masking)."""
    _, ids = model.full_sort_topk(interaction, k, mask_padding_item=True)
    pos_idx = (ids == positive_i.view(-1, 1).to(ids.dtype)).to(torch.int32)
    pos_len = torch.ones((ids.shape[0], 1), dtype=torch.int32, device=ids.device)
    return torch.cat([pos_idx, pos_len], dim=1)


def metrics_from_record(record, topk=(10, 20), decimal_place=4):
    """RecBole's metric definitions on a concatenated "rec.topk" record [n_users, K + 1] (one positive per user):
    hit@k = any(pos[:k]); mrr@k = 1 / rank of the first hit (0 if none); ndcg@k = sum_j pos_j / log2(j + 1) / IDCG, IDCG = 1;
    recall@k = hits / n_pos; precision@k = hits / k.  Means over users, rounded like RecBole (metric_decimal_place = 4)."""
    pos = record[:, :-1].to(torch.float64)
    n_pos = record[:, -1].to(torch.float64).clamp_min(1)
    K = pos.shape[1]
    rank = torch.arange(1, K + 1, dtype=torch.float64, device=record.device)
    out = {}
    for k in topk:
        assert k <= K, f"record holds top-{K} only"
        p = pos[:, :k]
        hits = p.sum(1)
        first = torch.where(p > 0, rank[:k], torch.full_like(p, float("inf"))).min(1).values
        out[f"hit@{k}"] = (hits > 0).double().mean()
        out[f"mrr@{k}"] = torch.where(torch.isinf(first), torch.zeros_like(first), 1.0 / first).mean()
        idcg = torch.cumsum(1.0 / torch.log2(rank + 1), 0)[(n_pos.clamp_max(k) - 1).long()]
        out[f"ndcg@{k}"] = ((p / torch.log2(rank[:k] + 1)).sum(1) / idcg).mean()
        out[f"recall@{k}"] = (hits / n_pos).mean()
        out[f"precision@{k}"] = (hits / k).mean()
    return {name: round(float(v), decimal_place) if decimal_place is not None else float(v) for name, v in out.items()}


@torch.no_grad()
def full_sort_evaluate(model, eval_batches, topk=(10, 20), item_field="item_id", decimal_place=4):
    """Drop-in for the body of RecBole's `Trainer.evaluate` on a FullSortEvalDataLoader of a sequential model:
    eval_batches yields interactions (dict-like of CUDA tensors incl. the held-out `item_field`).  Returns the metric
    dict RecBole would log (keys lower-cased `hit@10`, ...)."""
    was_training = model.training
    model.eval()
    records = [topk_record(model, inter, inter[item_field], max(topk)) for inter in eval_batches]
    model.train(was_training)
    return metrics_from_record(torch.cat(records, dim=0), topk, decimal_place)


@torch.no_grad()
def evaluate_unseen_users(model, item_id_lists, item_lengths, true_item_ids, k=10, batch_size=4096,
                          item_list_field="item_id_list", length_field="item_length"):
    """Batched form of the reference's per-user loop `evaluate_with_preprocessing` (run_with_unseen.py:196-261):
    that loop calls `full_sort_predict` with B = 1, copies the dense score row to the host, drops the pad column, and
    computes sklearn's `ndcg_score(k=10)` and an argpartition Hit@10 over a dense [users, n_items - 1] matrix.  Here the
    users go through the fused scorer `batch_size` at a time and only the [B, k] ids come back.

    item_id_lists [n, L] (right-padded with 0), item_lengths [n], true_item_ids [n] are integer tensors / arrays of item
    IDS (token -> id conversion stays with the caller); rows with true_item_ids <= 0 (unknown target, or a history that
    could not be converted: the reference `continue`s over those) are excluded from the means, as its `valid_rows` filter
    does.  One positive per user: ndcg@k = 1 / log2(rank + 1) if the target is ranked <= k else 0 (sklearn's value when no
    scores tie), hit@k = target in the top k.  Returns {'hit@k': float, 'ndcg@k': float} unrounded, like the reference."""
    dev = next(model.parameters()).device
    lists = torch.as_tensor(item_id_lists).long()
    lens = torch.as_tensor(item_lengths).long()
    true = torch.as_tensor(true_item_ids).long()
    keep = true > 0
    if not bool(keep.any()):
        return {f"hit@{k}": 0.0, f"ndcg@{k}": 0.0}
    lists, lens, true = lists[keep], lens[keep], true[keep]
    was_training = model.training
    model.eval()
    hit = torch.zeros((), dtype=torch.float64, device=dev)
    ndcg = torch.zeros((), dtype=torch.float64, device=dev)
    discount = 1.0 / torch.log2(torch.arange(2, k + 2, dtype=torch.float64, device=dev))
    for s in range(0, lists.shape[0], batch_size):
        inter = {item_list_field: lists[s:s + batch_size].to(dev, non_blocking=True),
                 length_field: lens[s:s + batch_size].to(dev, non_blocking=True)}
        _, ids = model.full_sort_topk(inter, k, mask_padding_item=True)
        pos = (ids == true[s:s + batch_size].to(dev).view(-1, 1)).to(torch.float64)
        hit += pos.sum()
        ndcg += (pos * discount).sum()
    model.train(was_training)
    n = lists.shape[0]
    return {f"hit@{k}": float(hit) / n, f"ndcg@{k}": float(ndcg) / n}
