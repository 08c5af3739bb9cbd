"""Offline stand-in for Whisper's tiktoken tokenizer.

The BPE vocabulary files of `openai-whisper` are not available offline, so synthetic
runs use this reversible byte-level vocabulary.  It exposes exactly the duck-typed
surface the hot path reads from a tokenizer -- `sot_sequence`, `eot`, `no_timestamps`,
`encode`, `decode`, `split_tokens_on_unicode`, `split_to_word_tokens` (reference
timing.py:102,105; retokenize.py:8,22,24,29) -- and any real Whisper tokenizer can be
passed to the alignment API instead.

ids 0..255 are raw bytes (a `char` unit is one token, as with the real vocabulary for
ASCII); ids 256..16639 are ASCII byte pairs standing in for `subword` pieces; special
ids follow the multilingual / English-only layouts of the published models.
"""
from __future__ import annotations

import string

PAIR_BASE = 256
PAIR_COUNT = 128 * 128
CJK_LIKE = frozenset({"zh", "ja", "th", "lo", "my", "yue"})


class ByteTokenizer:
    def __init__(self, multilingual: bool = True, language: str = "en", task: str = "transcribe", num_languages: int = 99):
        self.multilingual = multilingual
        self.language = language
        self.task = task
        self.num_languages = num_languages
        if multilingual:
            # <|endoftext|> <|sot|> [num_languages language tokens] <|translate|> <|transcribe|> <|startoflm|>
            # <|startofprev|> <|nospeech|> <|notimestamps|> <|0.00|> ...: the ids after the language block move with
            # its size (99 languages up to large-v2 -> notimestamps 50363; 100 for large-v3 -> 50364)
            extra = num_languages - 99
            self.eot, self.sot = 50257, 50258
            self.no_timestamps, self.timestamp_begin = 50363 + extra, 50364 + extra
            self.sot_sequence = (self.sot, 50259, 50359 + extra)  # <|sot|><|en|><|transcribe|>
        else:
            self.eot, self.sot = 50256, 50257
            self.no_timestamps, self.timestamp_begin = 50362, 50363
            self.sot_sequence = (self.sot,)

    # -- text -> ids ----------------------------------------------------------------
    def encode(self, text: str) -> list[int]:
        data = text.encode("utf-8")
        ids, pos = [], 0
        while pos < len(data):
            a = data[pos]
            b = data[pos + 1] if pos + 1 < len(data) else None
            # pair up ASCII bytes, but never let a piece swallow the space that opens the next word
            if b is not None and a < 128 and b < 128 and b != 0x20:
                ids.append(PAIR_BASE + a * 128 + b)
                pos += 2
            else:
                ids.append(a)
                pos += 1
        return ids

    # -- ids -> text ----------------------------------------------------------------
    @staticmethod
    def _payload(tok: int) -> bytes:
        if tok < PAIR_BASE:
            return bytes((tok,))
        if tok < PAIR_BASE + PAIR_COUNT:
            return bytes(divmod(tok - PAIR_BASE, 128))
        return b""

    def _special_text(self, tok: int) -> str:
        named = {self.eot: "<|endoftext|>", self.sot: "<|startoftranscript|>", self.no_timestamps: "<|notimestamps|>"}
        if tok in named:
            return named[tok]
        if tok >= self.timestamp_begin:
            return f"<|{(tok - self.timestamp_begin) * 0.02:.2f}|>"
        return f"<|special{tok}|>"

    def decode(self, tokens) -> str:
        raw = b"".join(self._payload(int(t)) for t in tokens if int(t) < self.eot)
        return raw.decode("utf-8", errors="replace")

    def decode_with_timestamps(self, tokens) -> str:
        parts, pending = [], bytearray()
        for t in map(int, tokens):
            if t < self.eot:
                pending += self._payload(t)
                continue
            if pending:
                parts.append(bytes(pending).decode("utf-8", errors="replace"))
                pending.clear()
            parts.append(self._special_text(t))
        if pending:
            parts.append(bytes(pending).decode("utf-8", errors="replace"))
        return "".join(parts)

    # -- grouping -------------------------------------------------------------------
    def split_tokens_on_unicode(self, tokens):
        """Cut the token stream wherever the text decoded so far is valid unicode."""
        whole = self.decode_with_timestamps(tokens)
        pieces, piece_tokens, open_tokens, consumed = [], [], [], 0
        for tok in tokens:
            open_tokens.append(tok)
            text = self.decode_with_timestamps(open_tokens)
            hole = text.find("�")
            if hole < 0 or whole[consumed + hole] == "�":
                pieces.append(text)
                piece_tokens.append(open_tokens)
                open_tokens = []
                consumed += len(text)
        return pieces, piece_tokens

    def split_tokens_on_spaces(self, tokens):
        pieces, piece_tokens = self.split_tokens_on_unicode(tokens)
        words, word_tokens = [], []
        for piece, toks in zip(pieces, piece_tokens):
            starts_word = (
                toks[0] >= self.eot or piece.startswith(" ") or piece.strip() in string.punctuation or not words
            )
            if starts_word:
                words.append(piece)
                word_tokens.append(toks)
            else:
                words[-1] += piece
                word_tokens[-1].extend(toks)
        return words, word_tokens

    def split_to_word_tokens(self, tokens):
        if self.language in CJK_LIKE:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)


def get_tokenizer(multilingual: bool = True, *, language: str | None = None, task: str | None = None,
                  num_languages: int = 99) -> ByteTokenizer:
    lang = (language or "en").lower()
    lang = {"english": "en"}.get(lang, lang)
    return ByteTokenizer(multilingual, lang, task or "transcribe", num_languages if multilingual else 0)
