"""RACE reading-comprehension dataset: every question becomes 4 (question+choice, article) sequences
(parity: tasks/race/data.py)."""
import glob
import json
import os
import time

from torch.utils.data import Dataset

from megatron_llm_b200 import print_rank_0
from tasks.data_utils import build_sample, build_tokens_types_paddings_from_ids, clean_text

NUM_CHOICES = 4
MAX_QA_LENGTH = 128


class RaceDataset(Dataset):
    def __init__(self, dataset_name, datapaths, tokenizer, max_seq_length, max_qa_length=MAX_QA_LENGTH):
        self.dataset_name = dataset_name
        print_rank_0(" > building RACE dataset for {}:".format(dataset_name))
        print_rank_0("  > paths: " + " ".join(datapaths))
        self.samples = []
        for path in datapaths:
            self.samples.extend(process_single_datapath(path, tokenizer, max_qa_length, max_seq_length))
        print_rank_0("  >> total number of samples: {}".format(len(self.samples)))
        self.sample_multiplier = NUM_CHOICES     # the batch the model sees is 4x the loader's batch

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx):
        return self.samples[idx]


def process_single_datapath(datapath, tokenizer, max_qa_length, max_seq_length):
    """``datapath``/*.txt, one JSON document per line: article, questions[], options[][4], answers[] ('A'..'D')."""
    print_rank_0("   > working on {}".format(datapath))
    t0 = time.time()
    samples = []
    n_docs = n_questions = 0
    for filename in sorted(glob.glob(os.path.join(datapath, "*.txt"))):
        with open(filename, "r") as f:
            for line in f:
                data = json.loads(line)
                n_docs += 1
                questions, choices, answers = data["questions"], data["options"], data["answers"]
                assert len(questions) == len(answers) == len(choices)
                context_ids = tokenizer.tokenize(clean_text(data["article"]))
                for qi, question in enumerate(questions):
                    n_questions += 1
                    label = ord(answers[qi]) - ord("A")
                    assert 0 <= label < NUM_CHOICES and len(choices[qi]) == NUM_CHOICES
                    ids_l, types_l, pads_l = [], [], []
                    for choice in choices[qi]:
                        # cloze questions carry a '_' placeholder, the others get the choice appended
                        qa = question.replace("_", choice) if "_" in question else " ".join([question, choice])
                        qa_ids = tokenizer.tokenize(clean_text(qa))[:max_qa_length]
                        ids, types, pads = build_tokens_types_paddings_from_ids(
                            qa_ids, context_ids, max_seq_length, tokenizer.cls, tokenizer.sep, tokenizer.pad)
                        ids_l.append(ids), types_l.append(types), pads_l.append(pads)
                    samples.append(build_sample(ids_l, types_l, pads_l, label, len(samples)))
    print_rank_0("    > processed {} document, {} questions, and {} samples in {:.2f} seconds".format(
        n_docs, n_questions, len(samples), time.time() - t0))
    return samples
