"""MSDP-EVAL-F1: F1 between generated and reference sentences (parity: tasks/msdp/evaluate.py)."""
from megatron_llm_b200 import get_args, print_rank_0
from tasks.msdp.metrics import F1Metric


def _read(path, clean):
    print_rank_0("reading %s" % path)
    with open(path, "r") as f:
        return [clean(line.strip()) for line in f]


def evaluate_f1(guess_file, answer_file):
    guesses = _read(guess_file, lambda s: s.replace("<|endoftext|>", ""))
    answers = _read(answer_file, lambda s: "" if s == "no_passages_used" else s)
    assert len(guesses) == len(answers), "lengths of guess and answer are different!"
    p, r, f1 = F1Metric.compute_all_pairs(guesses, answers)
    print_rank_0("Precision: %.4f; recall: %.4f; f1: %.4f" % (p, r, f1))
    print_rank_0("done :-)")
    return p, r, f1


def main():
    args = get_args()
    return evaluate_f1(args.guess_file, args.answer_file)
