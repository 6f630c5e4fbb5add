"""Word tokenisation for answer matching (parity: tasks/orqa/unsupervised/tokenizers.py, the DrQA/DPR tokenizers).

``SimpleTokenizer`` splits into alphanumeric runs and single non-space symbols and keeps character offsets; the spaCy
variant is available when spaCy is installed."""
import copy
import logging

import regex

logger = logging.getLogger(__name__)


class Tokens:
    """A tokenised text: rows of (text, text_with_whitespace, (start, end)[, pos, lemma, ner])."""
    TEXT, TEXT_WS, SPAN, POS, LEMMA, NER = range(6)

    def __init__(self, data, annotators, opts=None):
        self.data, self.annotators, self.opts = data, annotators, opts or {}

    def __len__(self):
        return len(self.data)

    def slice(self, i=None, j=None):
        out = copy.copy(self)
        out.data = self.data[i:j]
        return out

    def untokenize(self):
        return "".join(t[self.TEXT_WS] for t in self.data).strip()

    def words(self, uncased=False):
        return [t[self.TEXT].lower() if uncased else t[self.TEXT] for t in self.data]

    def offsets(self):
        return [t[self.SPAN] for t in self.data]

    def _column(self, name, idx):
        return [t[idx] for t in self.data] if name in self.annotators else None

    def pos(self):
        return self._column("pos", self.POS)

    def lemmas(self):
        return self._column("lemma", self.LEMMA)

    def entities(self):
        return self._column("ner", self.NER)

    def ngrams(self, n=1, uncased=False, filter_fn=None, as_strings=True):
        words = self.words(uncased)
        spans = [(s, e + 1) for s in range(len(words)) for e in range(s, min(s + n, len(words)))
                 if not (filter_fn and filter_fn(words[s:e + 1]))]
        return [" ".join(words[s:e]) for s, e in spans] if as_strings else spans

    def entity_groups(self):
        ents = self.entities()
        if not ents:
            return None
        non_ent = self.opts.get("non_ent", "O")
        groups, i = [], 0
        while i < len(ents):
            tag = ents[i]
            if tag != non_ent:
                j = i
                while j < len(ents) and ents[j] == tag:
                    j += 1
                groups.append((self.slice(i, j).untokenize(), tag))
                i = j
            else:
                i += 1
        return groups


class Tokenizer:
    def tokenize(self, text):
        raise NotImplementedError

    def shutdown(self):
        pass

    def __del__(self):
        self.shutdown()


class SimpleTokenizer(Tokenizer):
    ALPHA_NUM = r"[\p{L}\p{N}\p{M}]+"
    NON_WS = r"[^\p{Z}\p{C}]"

    def __init__(self, **kwargs):
        self._regexp = regex.compile("(%s)|(%s)" % (self.ALPHA_NUM, self.NON_WS),
                                     flags=regex.IGNORECASE + regex.UNICODE + regex.MULTILINE)
        if len(kwargs.get("annotators", {})) > 0:
            logger.warning("%s only tokenizes! Skipping annotators: %s", type(self).__name__, kwargs.get("annotators"))
        self.annotators = set()

    def tokenize(self, text):
        matches = list(self._regexp.finditer(text))
        data = []
        for i, m in enumerate(matches):
            start, end = m.span()
            ws_end = matches[i + 1].span()[0] if i + 1 < len(matches) else end
            data.append((m.group(), text[start:ws_end], (start, end)))
        return Tokens(data, self.annotators)


class SpacyTokenizer(Tokenizer):
    def __init__(self, **kwargs):
        import spacy
        self.annotators = copy.deepcopy(kwargs.get("annotators", set()))
        self.nlp = spacy.load(kwargs.get("model", "en_core_web_sm"))

    def tokenize(self, text):
        doc = self.nlp(text.replace("\n", " "))
        data = []
        for i, t in enumerate(doc):
            ws_end = doc[i + 1].idx if i + 1 < len(doc) else t.idx + len(t.text)
            data.append((t.text, text[t.idx:ws_end], (t.idx, t.idx + len(t.text)), t.tag_, t.lemma_, t.ent_type_))
        return Tokens(data, self.annotators, opts={"non_ent": ""})
