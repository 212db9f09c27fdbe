"""Word-level tokenizer of the CP ("content preservation") evaluator, importable on Python 3.12.

Restates the behaviour of /root/reference/evaluate/auto/tokenizer.py:137-142 (a DeepMoji-style
regex splitter: URLs, e-mails, hyphen/underscore compounds, #hashtags, @mentions, hearts,
emoticons, contractions, titles, abbreviations, numbers, words, runs of one symbol, emoji, any
other character; whitespace dropped).  The reference builds its pattern with an inline ``(?i)``
in the middle (tokenizer.py:37), which Python <= 3.10 applied to the WHOLE pattern and Python
>= 3.11 rejects; the reference therefore ran fully case-insensitive, and so does this one
(``re.IGNORECASE``).  Host-side text handling only: it feeds ``calculate_wmd_scores``.
"""
from __future__ import annotations

import re
from typing import List

_WORD = r"[a-zA-Z]+"
# the reference spells the pound sign as \xa3 inside a RAW string (tokenizer.py:42), i.e. as the four
# characters \ x a 3, so a pound sign is NOT a symbol run there (each one falls through to "any
# character"); 'x', 'a' and '3' runs are shadowed by the word / number classes.  Mirrored here.
_SYMBOL_CHARS = "()<!?.,/'\"-_=\\§|´ˇ°[]<>{}~$^&*;:%+€`"


def _emoticons() -> str:
    fixed = ["-_-", "x_x", "^_^", "o.o", "o_o", "(:", "):", ");", "(;"]
    alts = [re.escape(s) for s in fixed]
    for start in (">:", ":", "=", ";"):
        for mid in ("-", ",", "^", "'", '"'):
            for end in ("D", "d", "p", "P", "v", ")", "o", "O", "(", "3", "/", "|", "\\"):
                alts.append(f"{re.escape(start)}{re.escape(mid)}?{re.escape(end)}+")
    return "|".join(alts)


def _build() -> "re.Pattern[str]":
    symbol_runs = "|".join(re.escape(c) + "+" for c in _SYMBOL_CHARS)
    # '#'/'@' runs stop before a hashtag / mention: '##hello' -> '#', '#hello'
    symbol_runs += r"|#+(?=#[a-zA-Z0-9_]+)|@+(?=@[a-zA-Z0-9_]+)|#+|@+"
    token_classes = [
        r"(?:https?://|www\.)(?:[a-zA-Z]|[0-9]|[$-_@.&+]|[!*\(\),]|(?:%[0-9a-fA-F][0-9a-fA-F]))+",   # url
        r"\b[a-zA-Z0-9_.+-]+@[a-zA-Z0-9-]+\.[a-zA-Z0-9-.]+\b",                                       # e-mail
        r"[a-zA-Z]+[-_][a-zA-Z]+",                                                                    # compound
        r"#[a-zA-Z0-9_]+",                                                                            # hashtag
        r"@[a-zA-Z0-9_]+",                                                                            # mention
        r"(?:<+/?3+)+",                                                                               # heart
        _emoticons(),
        _WORD + r"'" + _WORD,                                                                         # contraction
        r"Mr\.|Ms\.|Mrs\.|Dr\.|Prof\.",                                                               # titles
        r"\b(?<!\.)(?:[A-Za-z]\.){2,}",                                                               # abbreviation
        r"[0-9]+",
        _WORD,
        symbol_runs,
        "\ud83c[\udf00-\udfff]|\ud83d[\udc00-\ude4f\ude80-\udeff]|[\u2600-\u26FF\u2700-\u27BF]",       # emoji
        r".",
    ]
    return re.compile(r"\s+|(" + "|".join(token_classes) + ")", re.UNICODE | re.IGNORECASE)


_PATTERN = _build()


def tokenize(text: str) -> List[str]:
    return [t for t in _PATTERN.findall(text) if t and t.strip()]
