"""Sentiment-classification data for the NLP distillation example (reference: example/distill/nlp/reader.py --
ChnSentiCorp TSV ``text_a<TAB>label`` read through a vocabulary into padded id sequences).

``TsvReader`` reads such files with a whitespace / character tokenizer and a frequency vocabulary;
``synthetic_corpus`` generates a learnable two-class corpus when no data is on disk (no network here)."""
import collections

import numpy as np

PAD, UNK = 0, 1


def tokenize(text: str):
    toks = text.strip().split()
    return toks if len(toks) > 1 else list(text.strip())     # Chinese text comes unsegmented: fall back to characters


class Vocab:
    def __init__(self, tokens=(), max_size=30000, min_freq=1):
        cnt = collections.Counter(tokens)
        self.itos = ["[PAD]", "[UNK]"] + [t for t, c in cnt.most_common(max_size - 2) if c >= min_freq]
        self.stoi = {t: i for i, t in enumerate(self.itos)}

    def __len__(self):
        return len(self.itos)

    def encode(self, text, seq_len):
        ids = [self.stoi.get(t, UNK) for t in tokenize(text)][:seq_len]
        return np.array(ids + [PAD] * (seq_len - len(ids)), dtype="int64")


class TsvReader:
    def __init__(self, path, vocab=None, seq_len=64, has_header=True):
        rows = []
        with open(path, encoding="utf-8") as f:
            for i, line in enumerate(f):
                if i == 0 and has_header:
                    continue
                parts = line.rstrip("\n").split("\t")
                if len(parts) >= 2:
                    rows.append((parts[0], int(parts[1])))
        self.rows, self.seq_len = rows, seq_len
        self.vocab = vocab or Vocab(t for text, _ in rows for t in tokenize(text))

    def samples(self):
        for text, label in self.rows:
            yield self.vocab.encode(text, self.seq_len), np.array([label], dtype="int64")


def synthetic_corpus(n, vocab, seq, seed):
    """Two classes that use different halves of the vocabulary, random lengths, zero padding."""
    rng = np.random.RandomState(seed)
    for _ in range(n):
        y = rng.randint(0, 2)
        ids = rng.randint(2, vocab // 2, size=seq) + (vocab // 2 - 2) * y
        ids[rng.randint(seq // 2, seq):] = PAD
        yield ids.astype("int64"), np.array([y], dtype="int64")


def batches(sample_iter, batch_size):
    buf = []
    for s in sample_iter:
        buf.append(s)
        if len(buf) == batch_size:
            yield np.stack([b[0] for b in buf]), np.stack([b[1] for b in buf])
            buf = []
