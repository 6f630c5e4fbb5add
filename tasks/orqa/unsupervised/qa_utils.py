"""Answer-presence checks for retrieved passages and reader predictions (parity: tasks/orqa/unsupervised/qa_utils.py)."""
import collections
import logging
import string
import unicodedata
from functools import partial
from multiprocessing import Pool as ProcessPool

import regex as re

from tasks.orqa.unsupervised.tokenizers import SimpleTokenizer

logger = logging.getLogger(__name__)
QAMatchStats = collections.namedtuple("QAMatchStats", ["top_k_hits", "questions_doc_hits"])
dpr_all_documents = None


def calculate_matches(all_docs, answers, closest_docs, workers_num, match_type):
    """``all_docs``: id -> (text, title); ``answers``: one list per question; ``closest_docs``: (ids, scores) per
    question.  Returns (top_k_hits[k] = #questions answered within the first k+1 passages, per-question hit lists)."""
    global dpr_all_documents
    dpr_all_documents = all_docs
    fn = partial(check_answer, match_type=match_type, tokenizer=SimpleTokenizer())
    pairs = list(zip(answers, closest_docs))
    if workers_num and workers_num > 1:
        with ProcessPool(processes=workers_num) as pool:      # fork: workers inherit ``dpr_all_documents``
            scores = pool.map(fn, pairs)
    else:
        scores = [fn(p) for p in pairs]
    n_docs = len(closest_docs[0][0])
    top_k_hits = [0] * n_docs
    for hits in scores:
        first = next((i for i, h in enumerate(hits) if h), None)
        if first is not None:
            for k in range(first, n_docs):
                top_k_hits[k] += 1
    return QAMatchStats(top_k_hits, scores)


def check_answer(questions_answers_docs, tokenizer, match_type):
    answers, (doc_ids, _) = questions_answers_docs
    hits = []
    for doc_id in doc_ids:
        text = dpr_all_documents[doc_id][0]
        if text is None:
            logger.warning("no doc in db")
            hits.append(False)
        else:
            hits.append(has_answer(answers, text, tokenizer, match_type))
    return hits


def has_answer(answers, text, tokenizer, match_type):
    """'string': token-sequence containment (uncased); 'regex': the answer is a pattern searched in the text."""
    text = _normalize(text)
    if match_type == "string":
        words = tokenizer.tokenize(text).words(uncased=True)
        for answer in answers:
            target = tokenizer.tokenize(_normalize(answer)).words(uncased=True)
            n = len(target)
            if any(target == words[i:i + n] for i in range(len(words) - n + 1)):
                return True
    elif match_type == "regex":
        return any(regex_match(text, _normalize(a)) for a in answers)
    return False


def regex_match(text, pattern):
    try:
        compiled = re.compile(pattern, flags=re.IGNORECASE + re.UNICODE + re.MULTILINE)
    except BaseException:
        return False
    return compiled.search(text) is not None


def exact_match_score(prediction, ground_truth):
    return _normalize_answer(prediction) == _normalize_answer(ground_truth)


def _normalize_answer(s):
    s = "".join(ch for ch in s.lower() if ch not in set(string.punctuation))
    s = re.sub(r"\b(a|an|the)\b", " ", s)
    return " ".join(s.split())


def _normalize(text):
    return unicodedata.normalize("NFD", text)
