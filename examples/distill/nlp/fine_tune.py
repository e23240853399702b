#!/usr/bin/env python
"""Fine-tune the sequence TEACHER and serve it (reference: example/distill/nlp/fine_tune.py fine-tunes ERNIE
through PaddleHub and exports it for Paddle Serving).  No pretrained checkpoint can be downloaded here, so the
teacher is a small Transformer encoder trained from scratch on the same corpus; the serving contract is the
same: feed ``ids`` [B, L] int64, fetch ``logits`` [B, 2].

    python examples/distill/nlp/fine_tune.py --epochs 3 --save teacher.pt
    python examples/distill/nlp/fine_tune.py --load teacher.pt --serve 9292     # then: distill.py --teachers 127.0.0.1:9292
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..", "..")))
from reader import batches, synthetic_corpus  # noqa: E402


class EncoderTeacher(nn.Module):
    def __init__(self, vocab, dim=128, heads=4, layers=2, seq_len=64, num_labels=2):
        super().__init__()
        self.emb = nn.Embedding(vocab, dim, padding_idx=0)
        self.pos = nn.Parameter(torch.zeros(seq_len, dim))
        layer = nn.TransformerEncoderLayer(dim, heads, dim * 4, dropout=0.1, batch_first=True)
        self.enc = nn.TransformerEncoder(layer, layers)
        self.head = nn.Linear(dim, num_labels)

    def forward(self, ids):
        ids = ids.long()
        pad = ids == 0
        h = self.enc(self.emb(ids) + self.pos[: ids.shape[1]], src_key_padding_mask=pad)
        keep = (~pad).unsqueeze(-1).float()
        return self.head((h * keep).sum(1) / keep.sum(1).clamp_min(1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--vocab", type=int, default=2000)
    ap.add_argument("--seq_len", type=int, default=64)
    ap.add_argument("--samples", type=int, default=1024)
    ap.add_argument("--save", default=None)
    ap.add_argument("--load", default=None)
    ap.add_argument("--serve", type=int, default=0, help="port: serve the teacher after training / loading")
    args = ap.parse_args()
    model = EncoderTeacher(args.vocab, seq_len=args.seq_len)
    if args.load:
        model.load_state_dict(torch.load(args.load, map_location="cpu"))
    else:
        opt = torch.optim.AdamW(model.parameters(), 5e-4)
        for ep in range(args.epochs):
            hit = n = 0
            for ids, label in batches(synthetic_corpus(args.samples, args.vocab, args.seq_len, ep), 16):
                logits = model(torch.from_numpy(ids))
                y = torch.from_numpy(label[:, 0])
                loss = torch.nn.functional.cross_entropy(logits, y)
                opt.zero_grad()
                loss.backward()
                opt.step()
                hit += int((logits.argmax(-1) == y).sum())
                n += len(y)
            print("epoch %d loss %.4f acc %.4f" % (ep, float(loss), hit / n), flush=True)
    if args.save:
        torch.save(model.state_dict(), args.save)
    if args.serve:
        from paddle_edl.distill.teacher_server import TeacherServer

        srv = TeacherServer(model.eval(), ["ids"], ["logits"], {"ids": [args.seq_len]}, port=args.serve).start()
        print("teacher serving on", srv.endpoint, flush=True)
        try:
            while True:
                time.sleep(3600)
        except KeyboardInterrupt:
            srv.stop()


if __name__ == "__main__":
    main()
