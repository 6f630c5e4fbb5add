"""MNLI (parity: tasks/glue/mnli.py).  Train/dev files have the gold label in the last column; the 10-column test file
has none and gets ``test_label``."""
from megatron_llm_b200 import print_rank_0
from tasks.data_utils import clean_text

from .data import GLUEAbstractDataset, read_tsv

LABELS = {"contradiction": 0, "entailment": 1, "neutral": 2}


class MNLIDataset(GLUEAbstractDataset):
    def __init__(self, name, datapaths, tokenizer, max_seq_length, test_label="contradiction"):
        self.test_label = test_label
        super().__init__("MNLI", name, datapaths, tokenizer, max_seq_length)

    def process_samples_from_single_path(self, filename):
        print_rank_0(" > Processing {} ...".format(filename))
        rows = read_tsv(filename)
        header = next(rows)
        is_test = len(header) == 10
        samples = []
        for row in rows:
            label = self.test_label if is_test else row[-1]
            sample = {"text_a": clean_text(row[8]), "text_b": clean_text(row[9]), "label": LABELS[label],
                      "uid": int(row[0])}
            assert sample["text_a"] and sample["text_b"] and sample["uid"] >= 0
            samples.append(sample)
        print_rank_0(" >> processed {} samples.".format(len(samples)))
        return samples
