"""Byte-pair-encoding tokenizer for the CLIP text encoders (host side of SURVEY 8f N2).

Restates the in-tree Swift tokenizer (``swift/StableDiffusion/tokenizer/BPETokenizer.swift:52-167`` and
``BPETokenizer+Reading.swift:15-50``): lower-case, trim, split on spaces, start from characters with
``</w>`` appended to the last one, repeatedly merge the adjacent pair with the lowest merge rank, map tokens to
ids with ``<|endoftext|>`` as the unknown token.  The padding / truncation to the encoder's input length follows
``TextEncoder.encode`` (``pipeline/TextEncoder.swift:52-68``); calling the object returns the ``(1, length)``
float32 ``input_ids`` array the Python reference feeds to the text encoder (``pipeline.py:151-175``).
Known-answer ids: ``StableDiffusionTests.swift:43-48`` (``tests/test_tokenizer.py``).
"""
from __future__ import annotations

import json

import numpy as np


class BPETokenizer:
    start_token = "<|startoftext|>"
    end_token = "<|endoftext|>"
    unknown_token = "<|endoftext|>"

    def __init__(self, merges: dict, vocabulary: dict, pad_token: str = "<|endoftext|>", model_max_length: int = 77):
        self.merges = merges              # {(first, second): rank}
        self.vocabulary = vocabulary      # {token: id}
        self.pad_token = pad_token
        self.model_max_length = model_max_length
        self._ids_to_tokens = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_files(cls, merges_path, vocabulary_path, pad_token="<|endoftext|>", model_max_length=77):
        """BPETokenizer(mergesAt:vocabularyAt:padToken:) (BPETokenizer.swift:46-50)."""
        return cls(cls.read_merges(merges_path), cls.read_vocabulary(vocabulary_path), pad_token, model_max_length)

    @staticmethod
    def read_vocabulary(path) -> dict:
        with open(path, "rb") as f:
            return {str(k): int(v) for k, v in json.loads(f.read().decode("utf-8")).items()}

    @staticmethod
    def read_merges(path) -> dict:
        """Rank = zero-based line number, counting blank and '#' lines too (BPETokenizer+Reading.swift:22-50).
        Like the Swift reader, a final line without a trailing newline is ignored."""
        merges = {}
        with open(path, "rb") as f:
            data = f.read()
        lines = data.split(b"\n")[:-1]
        for index, line in enumerate(lines):
            if not line or line[:1] == b"#":
                continue
            pair = [p for p in line.split(b" ") if p]
            if len(pair) != 2:
                raise ValueError(f"invalid merges file line {index + 1}")
            merges[(pair[0].decode("utf-8"), pair[1].decode("utf-8"))] = index
        return merges

    @property
    def unknown_token_id(self) -> int:
        return self.vocabulary.get(self.unknown_token, 0)

    # ------------------------------------------------------------------ encoding
    def tokenize(self, input: str, min_count: int | None = None):
        """-> (tokens, ids): start token, the BPE tokens of every word, end token, padded to ``min_count``."""
        tokens = [self.start_token] + self.encode(input) + [self.end_token]
        if min_count is not None and min_count > len(tokens):
            tokens += [self.pad_token] * (min_count - len(tokens))
        unk = self.unknown_token_id
        return tokens, [self.vocabulary.get(t, unk) for t in tokens]

    def token_id(self, token: str):
        return self.vocabulary.get(token)

    def token(self, id: int):
        if self._ids_to_tokens is None:
            self._ids_to_tokens = {v: k for k, v in self.vocabulary.items()}
        return self._ids_to_tokens.get(id)

    def decode(self, tokens) -> str:
        return "".join(tokens).replace("</w>", " ").replace(self.start_token, "").replace(self.end_token, "")

    def encode(self, input: str):
        words = [w for w in input.strip().lower().split(" ") if w]
        out = []
        for w in words:
            out += self.encode_word(w)
        return out

    def encode_word(self, word: str):
        tokens = list(word)
        if tokens:
            tokens[-1] += "</w>"
        while len(tokens) > 1:
            best, best_rank = None, None
            for pair in zip(tokens, tokens[1:]):
                rank = self.merges.get(pair)
                if rank is not None and (best_rank is None or rank < best_rank):
                    best, best_rank = pair, rank
            if best is None:
                break
            tokens = self._merge(tokens, best)
        return tokens

    @staticmethod
    def _merge(tokens, bigram):
        """Greedy left-to-right merge of every occurrence of ``bigram`` (BPETokenizer.swift:138-167)."""
        first, second = bigram
        out, i, n = [], 0, len(tokens)
        while i < n:
            if tokens[i] == first and i + 1 < n and tokens[i + 1] == second:
                out.append(first + second)
                i += 2
            else:
                out.append(tokens[i])
                i += 1
        return out

    # ------------------------------------------------------------------ encoder-facing call
    def input_ids(self, text: str, length: int | None = None):
        """Pad to / truncate at the text encoder's input length (TextEncoder.swift:52-68)."""
        length = length or self.model_max_length
        _, ids = self.tokenize(text, min_count=length)
        return ids[:length]

    def __call__(self, text: str):
        return np.asarray([self.input_ids(text)], dtype=np.float32)  # ids travel as float32 (pipeline.py:173)
