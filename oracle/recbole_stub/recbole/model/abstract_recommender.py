"""TEST INFRASTRUCTURE ONLY. Stand-in for recbole.model.abstract_recommender (RecBole 1.2.0,
[upstream], not installed): the attributes RecBLR.py:20,37-38,83,87-92 rely on."""
import torch
from torch import nn


class SequentialRecommender(nn.Module):
    def __init__(self, config, dataset):
        super().__init__()
        self.USER_ID = config["USER_ID_FIELD"]
        self.ITEM_ID = config["ITEM_ID_FIELD"]
        self.ITEM_SEQ = self.ITEM_ID + config["LIST_SUFFIX"]
        self.ITEM_SEQ_LEN = config["ITEM_LIST_LENGTH_FIELD"]
        self.POS_ITEM_ID = self.ITEM_ID
        self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
        self.max_seq_length = config["MAX_ITEM_LIST_LENGTH"]
        self.n_items = dataset.num(self.ITEM_ID)
        self.device = config["device"]

    def gather_indexes(self, output, gather_index):
        idx = gather_index.view(-1, 1, 1).expand(-1, -1, output.shape[-1])
        return output.gather(dim=1, index=idx).squeeze(1)
