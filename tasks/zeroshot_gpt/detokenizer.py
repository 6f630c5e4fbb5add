"""Undo the dataset-specific pre-tokenisation of PTB / WikiText / LAMBADA dumps before BPE tokenisation
(parity: tasks/zeroshot_gpt/detokenizer.py).  Table-driven: (literal | regex, replacement) pairs applied in order."""
import re

_PTB = [(" '", "'"), (" \n", "\n"), ("\n ", "\n"), (" n't", "n't"), (" N ", "1 "), ("$ 1", "$1"), ("# 1", "#1")]

_WIKI = [
    ("s '", "s'"), (re.compile(r"/' [0-9]/"), r"/'[0-9]/"),
    (" @-@ ", "-"), (" @,@ ", ","), (" @.@ ", "."),                       # number separators
    (" : ", ": "), (" ; ", "; "), (" . ", ". "), (" ! ", "! "), (" ? ", "? "), (" , ", ", "),   # punctuation
    (re.compile(r"\(\s*([^\)]*?)\s*\)"), r"(\1)"), (re.compile(r"\[\s*([^\]]*?)\s*\]"), r"[\1]"),
    (re.compile(r"{\s*([^}]*?)\s*}"), r"{\1}"), (re.compile(r"\"\s*([^\"]*?)\s*\""), r'"\1"'),
    (re.compile(r"'\s*([^']*?)\s*'"), r"'\1'"),                             # brackets / quotes
    ("= = = =", "===="), ("= = =", "==="), ("= =", "=="),                    # headings
    (" " + chr(176) + " ", chr(176)), (" \n", "\n"), ("\n ", "\n"), (" N ", " 1 "), (" 's", "'s"),
]


def _apply(rules, string):
    for pat, rep in rules:
        string = pat.sub(rep, string) if hasattr(pat, "sub") else string.replace(pat, rep)
    return string


def ptb_detokenizer(string):
    return _apply(_PTB, string)


def wikitext_detokenizer(string):
    return _apply(_WIKI, string)


def lambada_detokenizer(string):
    return string


_DETOKENIZERS = {"ptb": ptb_detokenizer, "wiki": wikitext_detokenizer, "lambada": lambada_detokenizer}


def get_detokenizer(path):
    for key, fn in _DETOKENIZERS.items():
        if key in path:
            return fn
    return None
