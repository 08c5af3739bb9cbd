"""ORACLE / TEST INFRASTRUCTURE ONLY: stub for the absent `num2words` package
(reference retokenize.py:2,46).  Spells small non-negative integers; enough for the
reference module to import and for digit-free synthetic text."""
_ONES = "zero one two three four five six seven eight nine ten eleven twelve thirteen fourteen fifteen sixteen seventeen eighteen nineteen".split()
_TENS = "_ _ twenty thirty forty fifty sixty seventy eighty ninety".split()


def num2words(n: int) -> str:
    n = int(n)
    if n < 20:
        return _ONES[n]
    if n < 100:
        return _TENS[n // 10] + ("" if n % 10 == 0 else "-" + _ONES[n % 10])
    if n < 1000:
        rest = n % 100
        return _ONES[n // 100] + " hundred" + ("" if rest == 0 else " and " + num2words(rest))
    rest = n % 1000
    return num2words(n // 1000) + " thousand" + ("" if rest == 0 else " " + num2words(rest))
