"""Offline csv -> id encoding: Tokenizer, Normalizer, pad_sequences.

Same contracts as the reference's fuxictr/preprocess.py (Tokenizer :30-129, Normalizer :142-163, pad_sequences :166-191):
ids start at 1 in (count descending, token ascending) order, 0 = __OOV__, the last id = __PAD__ when padding is used,
`min_freq` filters rare tokens, the `na_value` token never gets an id.  This is host-side data preparation with no GPU angle;
the implementation is array based (pandas.factorize + one dict look-up per DISTINCT token instead of per-row look-ups and a
process pool; sequences are padded / truncated with index arithmetic on the flattened token array).
"""
import json
import numpy as np
import pandas as pd


def _as_str_array(values):
    """object array of the raw cell values (csv columns are usually read with dtype=str; NaN stays NaN)"""
    return np.asarray(values, dtype=object)


def _count(tokens):
    """[(token, count)] of an object array; NaN cells are not tokens"""
    codes, uniques = pd.factorize(np.asarray(tokens, dtype=object))
    counts = np.bincount(codes[codes >= 0], minlength=len(uniques))
    return list(zip(list(uniques), counts.tolist()))


class Tokenizer(object):
    def __init__(self, num_words=None, na_value=None, min_freq=1, splitter=None, lower=False, oov_token=0, max_len=0,
                 padding="pre", num_workers=4):
        self._num_words = num_words
        self._na_value = na_value
        self._min_freq = min_freq
        self._lower = lower
        self._splitter = splitter
        self.oov_token = oov_token          # 0 = __OOV__
        self.vocab = dict()
        self.vocab_size = 0                 # includes oov and padding
        self.max_len = max_len
        self.padding = padding
        self.num_workers = num_workers      # accepted for signature parity; the array implementation is single pass
        self.use_padding = False

    # ------------------------------------------------------------------ fitting
    def _split_all(self, texts):
        """flat token array + per-row token counts of a sequence column (reference count_tokens :131-139: text.split(sep),
        so '' splits into [''] for an explicit separator and a NaN row raises -- rows are strings after fill_na)"""
        rows = [t.split(self._splitter) for t in texts]
        lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
        flat = np.asarray([tok for r in rows for tok in r], dtype=object) if len(rows) else np.empty(0, dtype=object)
        return flat, lens

    def fit_on_texts(self, texts, use_padding=False):
        self.use_padding = use_padding
        texts = _as_str_array(texts)
        if self._splitter is not None:                                   # sequence feature
            flat, lens = self._split_all(texts)
            if self.max_len == 0:                                        # argument max_len not given
                self.max_len = int(lens.max()) if len(lens) else 0
            word_counts = _count(flat)
        else:
            word_counts = _count(texts)
        self.build_vocab(word_counts)

    def build_vocab(self, word_counts):
        items = word_counts.items() if isinstance(word_counts, dict) else word_counts
        ordered = sorted(items, key=lambda kv: (-kv[1], kv[0]))          # deterministic index order
        words = [tok.lower() if self._lower else tok for tok, cnt in ordered
                 if cnt >= self._min_freq and (self._na_value is None or tok != self._na_value)]
        if self._num_words:
            words = words[0:self._num_words]
        self.vocab = dict((tok, idx) for idx, tok in enumerate(words, 1 + self.oov_token))
        self.vocab["__OOV__"] = self.oov_token
        if self.use_padding:
            self.vocab["__PAD__"] = len(words) + self.oov_token + 1      # the last index
        self.vocab_size = len(self.vocab) + self.oov_token

    def load_vocab_from_file(self, vocab_file):
        with open(vocab_file, "r") as fid:
            word_counts = json.load(fid)
        self.build_vocab(word_counts)

    def set_vocab(self, vocab):
        self.vocab = vocab
        self.vocab_size = len(self.vocab) + self.oov_token

    # ------------------------------------------------------------------ encoding
    def _lookup(self, tokens):
        """ids of an array of tokens (any hashable type); unknown tokens and NaN -> oov.  One dict look-up per DISTINCT token."""
        if len(tokens) == 0:
            return np.empty(0, dtype=np.int64)
        codes, uniques = pd.factorize(np.asarray(tokens, dtype=object))
        ids = np.fromiter((self.vocab.get(t, self.oov_token) for t in uniques), dtype=np.int64, count=len(uniques))
        ids = np.append(ids, self.oov_token)                            # code -1 (NaN) -> the appended oov slot
        return ids[codes]

    def encode_category(self, categories):
        return self._lookup(_as_str_array(categories))

    def encode_sequence(self, texts):
        texts = _as_str_array(texts)
        n = len(texts)
        empty = np.fromiter((pd.isnull(t) or t == "" for t in texts), dtype=bool, count=n)
        rows = [[] if e else t.split(self._splitter) for t, e in zip(texts, empty)]
        lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=n)
        flat = np.asarray([tok for r in rows for tok in r], dtype=object)
        ids = self._lookup(flat)
        out = np.full((n, self.max_len), self.vocab_size - 1, dtype=np.int32)
        if len(ids) == 0 or self.max_len == 0:
            return out
        # position of every token inside its row, then the reference's pad_sequences geometry (padding == truncating side)
        row_of = np.repeat(np.arange(n), lens)
        pos = np.arange(len(ids)) - np.repeat(np.cumsum(lens) - lens, lens)
        L = lens[row_of]
        if self.padding == "pre":                      # keep the LAST max_len tokens, right aligned
            col = self.max_len - L + pos
            keep = pos >= L - self.max_len
        else:                                          # keep the FIRST max_len tokens, left aligned
            col = pos
            keep = pos < self.max_len
        out[row_of[keep], col[keep]] = ids[keep]
        return out

    def load_pretrained_embedding(self, *args, **kwargs):
        raise NotImplementedError("pretrained embeddings are not supported by the B200 RAT path (no RAT configuration uses "
                                  "them; the model rejects `pretrained_emb` feature specs as well)")


class Normalizer(object):
    """numeric features: 'StandardScaler' / 'MinMaxScaler' (population statistics, as sklearn) or any callable"""

    def __init__(self, normalizer):
        self.callable = callable(normalizer)
        if not self.callable and normalizer not in ("StandardScaler", "MinMaxScaler"):
            raise NotImplementedError("normalizer={}".format(normalizer))
        self.normalizer = normalizer
        self._a = self._b = None

    def fit(self, X):
        if self.callable:
            return
        X = np.asarray(X, dtype=np.float64).reshape(-1)
        if self.normalizer == "StandardScaler":
            mean, std = np.nanmean(X), np.nanstd(X)
            self._a, self._b = mean, (std if std != 0 else 1.0)          # sklearn leaves constant columns unscaled
        else:
            lo, hi = np.nanmin(X), np.nanmax(X)
            self._a, self._b = lo, ((hi - lo) if hi != lo else 1.0)

    def normalize(self, X):
        if self.callable:
            return self.normalizer(X)
        return (np.asarray(X, dtype=np.float64).reshape(-1) - self._a) / self._b


def pad_sequences(sequences, maxlen=None, dtype="int32", padding="pre", truncating="pre", value=0.):
    """list of lists -> [n, maxlen] array (the tf.keras.preprocessing.sequence.pad_sequences contract, reference :166-191)"""
    assert padding in ["pre", "post"], "Invalid padding={}.".format(padding)
    assert truncating in ["pre", "post"], "Invalid truncating={}.".format(truncating)
    if maxlen is None:
        maxlen = max(len(x) for x in sequences)
    arr = np.full((len(sequences), maxlen), value, dtype=dtype)
    for idx, x in enumerate(sequences):
        if len(x) == 0:
            continue
        trunc = np.asarray(x[-maxlen:] if truncating == "pre" else x[:maxlen], dtype=dtype)
        if padding == "pre":
            arr[idx, -len(trunc):] = trunc
        else:
            arr[idx, :len(trunc)] = trunc
    return arr
