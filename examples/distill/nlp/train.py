#!/usr/bin/env python
"""Train the small student (BOW / TextCNN) alone on hard labels -- the baseline the distilled student is
compared with (reference: example/distill/nlp/train.py)."""
import argparse
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..", "..")))
from reader import TsvReader, batches, synthetic_corpus  # noqa: E402
from paddle_edl.models.small import BOW, TextCNN  # noqa: E402


def evaluate(model, data):
    model.eval()
    hit = n = 0
    with torch.no_grad():
        for ids, label in data:
            pred = model(torch.from_numpy(ids)).argmax(-1)
            hit += int((pred == torch.from_numpy(label[:, 0])).sum())
            n += len(label)
    model.train()
    return hit / max(1, n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="bow", choices=["bow", "cnn"])
    ap.add_argument("--train_tsv", default=None)
    ap.add_argument("--dev_tsv", default=None)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--vocab", type=int, default=2000)
    ap.add_argument("--seq_len", type=int, default=64)
    ap.add_argument("--samples", type=int, default=1024)
    args = ap.parse_args()
    if args.train_tsv:
        tr = TsvReader(args.train_tsv, seq_len=args.seq_len)
        dev = TsvReader(args.dev_tsv or args.train_tsv, vocab=tr.vocab, seq_len=args.seq_len)
        vocab = len(tr.vocab)
        train_iter = lambda ep: batches(tr.samples(), 16)       # noqa: E731
        dev_data = list(batches(dev.samples(), 16))
    else:
        vocab = args.vocab
        train_iter = lambda ep: batches(synthetic_corpus(args.samples, vocab, args.seq_len, ep), 16)   # noqa: E731
        dev_data = list(batches(synthetic_corpus(256, vocab, args.seq_len, 10_000), 16))
    model = BOW(vocab) if args.model == "bow" else TextCNN(vocab)
    opt = torch.optim.AdamW(model.parameters(), 1e-3)
    for ep in range(args.epochs):
        for ids, label in train_iter(ep):
            loss = torch.nn.functional.cross_entropy(model(torch.from_numpy(ids)), torch.from_numpy(label[:, 0]))
            opt.zero_grad()
            loss.backward()
            opt.step()
        print("epoch %d loss %.4f dev acc %.4f" % (ep, float(loss), evaluate(model, dev_data)), flush=True)


if __name__ == "__main__":
    main()
