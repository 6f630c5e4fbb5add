"""Token-level precision / recall / F1 for dialogue evaluation (parity: tasks/msdp/metrics.py)."""
import re
from collections import Counter
from typing import List

import numpy as np

re_art = re.compile(r"\b(a|an|the)\b")
re_punc = re.compile(r'[!"#$%&()*+,-./:;<=>?@\[\]\\^`{|}~_\']')


def normalize_answer(s):
    """lower-case, strip punctuation and articles, collapse whitespace."""
    return " ".join(re_art.sub(" ", re_punc.sub(" ", s.lower())).split())


class F1Metric:
    @staticmethod
    def _prec_recall_f1_score(pred_items, gold_items):
        same = sum((Counter(gold_items) & Counter(pred_items)).values())
        if same == 0:
            return 0, 0, 0
        p, r = same / len(pred_items), same / len(gold_items)
        return p, r, 2 * p * r / (p + r)

    @staticmethod
    def compute_each_pair(guess: str, answer: str):
        if answer == "":
            return None, None, None         # unanswerable reference: skipped in the average
        if guess == "":
            return 0, 0, 0
        return F1Metric._prec_recall_f1_score(normalize_answer(guess).split(), normalize_answer(answer).split())

    @staticmethod
    def compute_all_pairs(guesses: List[str], answers: List[str]):
        assert len(guesses) == len(answers)
        rows = [F1Metric.compute_each_pair(g, a) for g, a in zip(guesses, answers)]
        rows = [r for r in rows if r[0] is not None]
        p, r, f = zip(*rows) if rows else ((0,), (0,), (0,))
        return np.mean(p), np.mean(r), np.mean(f)
