"""GLUE sentence-pair dataset base (parity: tasks/glue/data.py)."""
from abc import ABC, abstractmethod

from torch.utils.data import Dataset

from megatron_llm_b200 import print_rank_0
from tasks.data_utils import build_sample, build_tokens_types_paddings_from_text


class GLUEAbstractDataset(ABC, Dataset):
    def __init__(self, task_name, dataset_name, datapaths, tokenizer, max_seq_length):
        self.task_name, self.dataset_name = task_name, dataset_name
        self.tokenizer, self.max_seq_length = tokenizer, max_seq_length
        print_rank_0(" > building {} dataset for {}:".format(task_name, dataset_name))
        print_rank_0("  > paths: " + " ".join(datapaths))
        self.samples = []
        for path in datapaths:
            self.samples.extend(self.process_samples_from_single_path(path))
        print_rank_0("  >> total number of samples: {}".format(len(self.samples)))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        raw = self.samples[idx]
        ids, types, paddings = build_tokens_types_paddings_from_text(raw["text_a"], raw["text_b"], self.tokenizer,
                                                                     self.max_seq_length)
        return build_sample(ids, types, paddings, raw["label"], raw["uid"])

    @abstractmethod
    def process_samples_from_single_path(self, datapath):
        """file -> list of {'text_a': str, 'text_b': str, 'label': int, 'uid': int}."""


def read_tsv(filename):
    """Rows of a tab-separated file (header first)."""
    with open(filename, "r") as f:
        for line in f:
            yield [c.strip() for c in line.strip().split("\t")]
