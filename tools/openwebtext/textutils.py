"""Dependency-free text helpers shared by the openwebtext tools.

The reference scripts import ftfy / langdetect / tldextract / LSH, none of which exist on an air-gapped box; each helper
prefers the library when it is importable and otherwise falls back to a self-contained implementation."""
import json
import re
import unicodedata
from urllib.parse import urlparse

_CONTROL = re.compile(r"[\x00-\x08\x0b\x0c\x0e-\x1f\x7f]")
_MOJIBAKE = {"â€™": "’", "â€œ": "“", "â€\x9d": "”", "â€“": "–", "â€”": "—", "Ã©": "é", "Ã¨": "è", "Ã¶": "ö",
             "Ã¼": "ü", "Ã¤": "ä", "Â ": " "}
_STOPWORDS = frozenset("the of and to in a is that for it as was with be by on not he this are or his from at which "
                       "but have an they you were her all she there would their we him been has when who will more if "
                       "no out so said what up its about into than them can only other".split())


def fix_text(text):
    try:
        import ftfy
        return ftfy.fix_text(text)
    except ImportError:
        pass
    for bad, good in _MOJIBAKE.items():
        if bad in text:
            text = text.replace(bad, good)
    return _CONTROL.sub("", unicodedata.normalize("NFC", text))


def detect_language(text):
    """'en' or 'other' (langdetect's code when it is installed)."""
    try:
        from langdetect import detect
        return detect(text)
    except ImportError:
        pass
    words = re.findall(r"[A-Za-z']+", text.lower())
    if not words:
        return "other"
    letters = sum(c.isalpha() for c in text)
    ascii_letters = sum(c.isascii() and c.isalpha() for c in text)
    stop = sum(w in _STOPWORDS for w in words) / len(words)
    return "en" if letters and ascii_letters / letters > 0.9 and stop > 0.08 else "other"


def registered_domain(url):
    try:
        import tldextract
        ext = tldextract.extract(url)
        return ext.domain, ext.suffix
    except ImportError:
        host = (urlparse(url).hostname or "").lower()
        parts = host.split(".")
        if len(parts) < 2:
            return host, ""
        two_level = len(parts) >= 3 and parts[-2] in ("co", "com", "org", "net", "ac", "gov", "edu") and len(parts[-1]) == 2
        return (parts[-3], ".".join(parts[-2:])) if two_level else (parts[-2], parts[-1])


def read_jsonl(path):
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            if line.strip():
                yield json.loads(line)


def write_jsonl_row(f, obj):
    f.write(json.dumps(obj, ensure_ascii=False).encode("utf-8") + b"\n")
