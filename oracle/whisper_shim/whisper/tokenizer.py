"""ORACLE / TEST INFRASTRUCTURE ONLY.

Stand-in for `openai-whisper`'s tiktoken-backed tokenizer (third party; the BPE
vocabulary files are not available offline).  It restates the *interface* the
reference's hot path touches -- sot_sequence, eot, no_timestamps, encode, decode,
split_tokens_on_unicode, split_tokens_on_spaces, split_to_word_tokens
(timing.py:102,105; retokenize.py:8,22,24,29) -- over a reversible byte-level
vocabulary: ids 0..255 are raw bytes (so a `char` unit is one token, as with the
real vocabulary for ASCII), ids 256..16639 are ASCII byte pairs (a deterministic
stand-in for `subword` units).  Special ids follow the multilingual layout.
"""
from __future__ import annotations

import string
from dataclasses import dataclass

_PAIR_BASE = 256
_N_PAIR = 128 * 128


@dataclass
class Tokenizer:
    multilingual: bool = True
    language: str = "en"
    task: str = "transcribe"

    def __post_init__(self):
        if self.multilingual:
            self.eot, self.sot = 50257, 50258
            self._lang, self._task = 50259, 50359
            self.no_timestamps, self.timestamp_begin = 50363, 50364
            self.sot_sequence = (self.sot, self._lang, self._task)
        else:
            self.eot, self.sot = 50256, 50257
            self.no_timestamps, self.timestamp_begin = 50362, 50363
            self.sot_sequence = (self.sot,)

    # ---- text <-> ids -------------------------------------------------
    def encode(self, text: str):
        """Greedy ASCII byte-pair pieces inside each space-prefixed word."""
        ids = []
        raw = text.encode("utf-8")
        i, n = 0, len(raw)
        while i < n:
            # a piece never spans a word start: a space may only open a piece
            if i + 1 < n and raw[i] < 128 and raw[i + 1] < 128 and raw[i + 1] != 0x20:
                ids.append(_PAIR_BASE + raw[i] * 128 + raw[i + 1])
                i += 2
            else:
                ids.append(raw[i])
                i += 1
        return ids

    def _bytes_of(self, tok: int) -> bytes:
        if tok < 256:
            return bytes([tok])
        if tok < _PAIR_BASE + _N_PAIR:
            t = tok - _PAIR_BASE
            return bytes([t // 128, t % 128])
        return b""

    def _special(self, tok: int) -> str:
        if tok == self.eot:
            return "<|endoftext|>"
        if tok == self.sot:
            return "<|startoftranscript|>"
        if tok == self.no_timestamps:
            return "<|notimestamps|>"
        if tok >= self.timestamp_begin:
            return f"<|{(tok - self.timestamp_begin) * 0.02:.2f}|>"
        return f"<|special{tok}|>"

    def decode(self, tokens) -> str:
        data = b"".join(self._bytes_of(int(t)) for t in tokens if int(t) < self.eot)
        return data.decode("utf-8", errors="replace")

    def decode_with_timestamps(self, tokens) -> str:
        out, run = [], b""
        for t in tokens:
            t = int(t)
            if t >= self.eot:
                if run:
                    out.append(run.decode("utf-8", errors="replace"))
                    run = b""
                out.append(self._special(t))
            else:
                run += self._bytes_of(t)
        if run:
            out.append(run.decode("utf-8", errors="replace"))
        return "".join(out)

    # ---- word grouping ------------------------------------------------
    def split_tokens_on_unicode(self, tokens):
        full = self.decode_with_timestamps(tokens)
        bad = "�"
        words, word_tokens, cur, off = [], [], [], 0
        for tok in tokens:
            cur.append(tok)
            dec = self.decode_with_timestamps(cur)
            if bad not in dec or full[off + dec.index(bad)] == bad:
                words.append(dec)
                word_tokens.append(cur)
                cur = []
                off += len(dec)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens):
        subwords, subword_tokens = self.split_tokens_on_unicode(tokens)
        words, word_tokens = [], []
        for sw, toks in zip(subwords, subword_tokens):
            special = toks[0] >= self.eot
            with_space = sw.startswith(" ")
            punct = sw.strip() in string.punctuation
            if special or with_space or punct or not words:
                words.append(sw)
                word_tokens.append(toks)
            else:
                words[-1] = words[-1] + sw
                word_tokens[-1].extend(toks)
        return words, word_tokens

    def split_to_word_tokens(self, tokens):
        if self.language in {"zh", "ja", "th", "lo", "my", "yue"}:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)


def get_tokenizer(multilingual: bool, *, num_languages: int = 99, language=None, task=None):
    lang = {"english": "en"}.get((language or "en").lower(), (language or "en").lower())
    return Tokenizer(multilingual=multilingual, language=lang, task=task or "transcribe")
