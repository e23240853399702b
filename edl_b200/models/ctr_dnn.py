"""CTR-DNN (Criteo click-through model) -- the reference's CTR example network:
26 sparse-slot embeddings ([1 000 001, 10] each, average-pooled per slot) + 13 dense features
-> concat(273) -> FC400 x3 (ReLU) -> FC2 -> softmax, trained with Adam 1e-4, AUC metric
(example/ctr/ctr/save_program.py:75-144, train.py:234).  The reference runs it CPU-only in
parameter-server mode; here it is an elastic data-parallel GPU model whose embedding gradients are
all-reduced densely (the BASELINE "embedding all-reduce bandwidth sweep" config)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class CtrDnn(nn.Module):
    def __init__(self, sparse_feature_dim=1000001, embedding_size=10, num_sparse=26, num_dense=13,
                 hidden=(400, 400, 400), shared_table=False):
        super().__init__()
        self.num_sparse, self.num_dense = num_sparse, num_dense
        n_tables = 1 if shared_table else num_sparse
        self.tables = nn.ModuleList([nn.EmbeddingBag(sparse_feature_dim, embedding_size, mode="mean", sparse=False)
                                     for _ in range(n_tables)])
        for t in self.tables:
            nn.init.uniform_(t.weight, -1.0 / math.sqrt(sparse_feature_dim), 1.0 / math.sqrt(sparse_feature_dim))
        dims = [num_sparse * embedding_size + num_dense] + list(hidden)
        self.fcs = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        for fc in self.fcs:
            nn.init.normal_(fc.weight, 0.0, 1.0 / math.sqrt(fc.in_features))
        self.out = nn.Linear(dims[-1], 2)

    def forward(self, dense, sparse_ids, sparse_offsets=None):
        """dense [B, 13] float; sparse_ids [B, 26, L] int64 (L ids per slot, average pooled) or, with
        ``sparse_offsets``, a list of 26 (ids, offsets) bags."""
        embs = []
        for s in range(self.num_sparse):
            table = self.tables[s % len(self.tables)]
            if sparse_offsets is None:
                embs.append(table(sparse_ids[:, s]))
            else:
                embs.append(table(sparse_ids[s], sparse_offsets[s]))
        x = torch.cat(embs + [dense], dim=1)
        for fc in self.fcs:
            x = F.relu(fc(x))
        return self.out(x)


def auc(scores, labels, num_thresholds=4096):
    """Streaming-friendly AUC from score histograms (what the ``auc`` op of the reference computes)."""
    s = scores.detach().float().clamp(0, 1)
    idx = (s * (num_thresholds - 1)).long()
    pos = torch.zeros(num_thresholds, device=s.device).index_add_(0, idx, labels.float())
    neg = torch.zeros(num_thresholds, device=s.device).index_add_(0, idx, 1.0 - labels.float())
    tp = torch.flip(torch.cumsum(torch.flip(pos, [0]), 0), [0])
    fp = torch.flip(torch.cumsum(torch.flip(neg, [0]), 0), [0])
    tot_p, tot_n = pos.sum().clamp_min(1), neg.sum().clamp_min(1)
    tpr = torch.cat([tp / tot_p, tp.new_zeros(1)])
    fpr = torch.cat([fp / tot_n, fp.new_zeros(1)])
    return torch.trapz(torch.flip(tpr, [0]), torch.flip(fpr, [0]))
