"""TEST INFRASTRUCTURE ONLY. Stand-in for recbole.model.loss.BPRLoss ([upstream])."""
import torch
from torch import nn


class BPRLoss(nn.Module):
    def __init__(self, gamma=1e-10):
        super().__init__()
        self.gamma = gamma

    def forward(self, pos_score, neg_score):
        return -torch.log(self.gamma + torch.sigmoid(pos_score - neg_score)).mean()
