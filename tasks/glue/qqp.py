"""QQP (parity: tasks/glue/qqp.py).  6-column train/dev (id, qid1, qid2, q1, q2, is_duplicate) or 3-column test."""
from megatron_llm_b200 import print_rank_0
from tasks.data_utils import clean_text

from .data import GLUEAbstractDataset, read_tsv

LABELS = [0, 1]


class QQPDataset(GLUEAbstractDataset):
    def __init__(self, name, datapaths, tokenizer, max_seq_length, test_label=0):
        self.test_label = test_label
        super().__init__("QQP", name, datapaths, tokenizer, max_seq_length)

    def process_samples_from_single_path(self, filename):
        print_rank_0(" > Processing {} ...".format(filename))
        rows = read_tsv(filename)
        header = next(rows)
        is_test = len(header) == 3
        assert is_test or len(header) == 6
        samples = []
        for row in rows:
            if is_test:
                assert len(row) == 3, "expected length 3: {}".format(row)
                uid, a, b, label = int(row[0]), clean_text(row[1]), clean_text(row[2]), self.test_label
                assert a and b
            else:
                if len(row) != 6:
                    print_rank_0("***WARNING*** index error, skipping: {}".format(row))
                    continue
                uid, a, b, label = int(row[0]), clean_text(row[3]), clean_text(row[4]), int(row[5])
                if not a or not b:
                    print_rank_0("***WARNING*** zero length question, skipping: {}".format(row))
                    continue
            assert label in LABELS and uid >= 0
            samples.append({"uid": uid, "text_a": a, "text_b": b, "label": label})
        print_rank_0(" >> processed {} samples.".format(len(samples)))
        return samples
