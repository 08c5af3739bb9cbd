"""Text <-> alignment-unit helpers; same names, arguments and behaviour as the
reference's retokenize.py (encode :5-17, split_tokens_on_spaces :19-39,
remove_punctuation :41-50).  Pure host code: it defines the word boundaries that the
device-side boundary extraction indexes with."""
from __future__ import annotations

import string

_UNITS = ("char", "subword")


def encode(text, tokenizer, aligned_unit_type="subword"):
    """Token ids of `text`.  `char`: every character encoded on its own, one space token
    between words (leading/trailing/multiple blanks collapse, as str.split does)."""
    assert aligned_unit_type in _UNITS
    if aligned_unit_type == "subword":
        return tokenizer.encode(text)
    blank = tokenizer.encode(" ")
    out = []
    for n, word in enumerate(text.split()):
        if n:
            out.extend(blank)
        for ch in word:
            out.extend(tokenizer.encode(ch))
    return out


def split_tokens_on_spaces(tokens, tokenizer, aligned_unit_type="subword"):
    """(words, word_tokens).  `char`: a word opens at the first piece, at a piece that is
    exactly one space, and at special tokens (>= eot); everything else extends the word."""
    assert aligned_unit_type in _UNITS
    if aligned_unit_type == "subword":
        return tokenizer.split_to_word_tokens(tokens)
    pieces, piece_tokens = tokenizer.split_tokens_on_unicode(tokens)
    words, word_tokens = [], []
    for piece, toks in zip(pieces, piece_tokens):
        if not words or piece == " " or toks[0] >= tokenizer.eot:
            words.append(piece)
            word_tokens.append(toks)
        else:
            words[-1] = words[-1] + piece
            word_tokens[-1].extend(toks)
    return words, word_tokens


_KEEP_APOSTROPHE = str.maketrans("", "", string.punctuation.replace("'", ""))


def _spell_number(n: int) -> str:
    try:
        from num2words import num2words  # optional dependency of the reference
        return num2words(n)
    except ImportError:
        ones = ("zero one two three four five six seven eight nine ten eleven twelve thirteen fourteen "
                "fifteen sixteen seventeen eighteen nineteen").split()
        tens = "_ _ twenty thirty forty fifty sixty seventy eighty ninety".split()
        if n < 20:
            return ones[n]
        if n < 100:
            return tens[n // 10] + ("" if n % 10 == 0 else "-" + ones[n % 10])
        if n < 1000:
            return ones[n // 100] + " hundred" + ("" if n % 100 == 0 else " and " + _spell_number(n % 100))
        head, rest = divmod(n, 1000)
        return _spell_number(head) + " thousand" + ("" if rest == 0 else " " + _spell_number(rest))


def remove_punctuation(text):
    """Drop punctuation except apostrophes and spell out all-digit words."""
    text = text.translate(_KEEP_APOSTROPHE)
    cleaned = []
    for word in text.split():
        if word.isdigit():
            word = _spell_number(int(word))
        cleaned.append(word.strip(string.punctuation))
    return " ".join(cleaned).translate(_KEEP_APOSTROPHE)
