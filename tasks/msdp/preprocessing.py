"""Data preparation for multi-stage dialogue prompting (parity: tasks/msdp/preprocessing.py).

``--func`` selects: process_wow_dataset | process_woi_dataset (raw dumps -> "topic \\t context \\t knowledge \\t response"),
get_knwl_gen_prompts (nearest-dialogue prompt selection with a DPR question encoder), get_resp_gen_prompts,
prepare_input."""
import argparse
import json

import numpy as np
import torch

from tasks.msdp.prompt import word_tokenize

NO_PASSAGE = "no_passages_used"


def get_args():
    p = argparse.ArgumentParser(description="Preprocessing")
    p.add_argument("--func", type=str, default=None, help="choose to run which function")
    p.add_argument("--raw_file", type=str, default=None, help="path of the input file")
    p.add_argument("--processed_file", type=str, default=None, help="path of the output file")
    p.add_argument("--knwl_ref_file", type=str, default=None, help="path of the knowledge reference file")
    p.add_argument("--resp_ref_file", type=str, default=None, help="path of the response reference file")
    p.add_argument("--knwl_gen_file", type=str, default=None, help="path of the generated knowledge file")
    p.add_argument("--test_file", type=str, default=None, help="path of the test file")
    p.add_argument("--train_file", type=str, default=None, help="path of the train file")
    p.add_argument("--model_file", type=str, default=None, help="path of the DPR question encoder")
    p.add_argument("--data_type", type=str, default=None, help="wow_seen | wow_unseen | woi")
    p.add_argument("--seed", type=int, default=1234, help="random seed")
    return p.parse_args()


class _Writers:
    """processed file + optional knowledge / response reference files."""

    def __init__(self, processed_file, knwl_ref_file, resp_ref_file):
        self.proc = open(processed_file, "w")
        self.knwl = open(knwl_ref_file, "w") if knwl_ref_file else None
        self.resp = open(resp_ref_file, "w") if resp_ref_file else None

    def emit(self, topic, context, knowledge, response):
        self.proc.write("\t".join([topic, context, knowledge, response]) + "\n")
        if self.knwl:
            self.knwl.write(knowledge + "\n")
        if self.resp:
            self.resp.write(" ".join(word_tokenize(response)) + "\n")

    def close(self):
        for f in (self.proc, self.knwl, self.resp):
            if f:
                f.close()


def process_wow_dataset(raw_file, processed_file, knwl_ref_file, resp_ref_file):
    """Wizard of Wikipedia: every wizard turn after the first becomes one sample."""
    with open(raw_file, "r") as f:
        dialogs = json.load(f)
    out = _Writers(processed_file, knwl_ref_file, resp_ref_file)
    for sample in dialogs:
        history = []
        for j, turn in enumerate(sample["dialog"]):
            text = turn["text"]
            if not text.endswith(("?", ".", "!")):
                text += "."
            if j == 0 or "wizard" not in turn["speaker"].lower():
                assert j == 0 or "apprentice" in turn["speaker"].lower()
                history.append(text)
                continue
            sentences = list(turn["checked_sentence"].values())
            passages = list(turn["checked_passage"].values())
            assert len(sentences) <= 1
            knowledge = sentences[0] if sentences else NO_PASSAGE
            passage = passages[0] if len(passages) == 1 else NO_PASSAGE
            topic = passage if passage != NO_PASSAGE else sample["chosen_topic"]
            out.emit(topic, " [SEP] ".join(history), knowledge, text)
            history.append(text)
    out.close()


def process_woi_dataset(raw_file, processed_file, knwl_ref_file, resp_ref_file):
    """Wizard of Internet (json lines): topic = the wizard's last search query, knowledge = the selected sentence."""
    strip = lambda s: s.replace("\n", "").replace("\r", "").replace("\t", "")      # noqa: E731
    out = _Writers(processed_file, knwl_ref_file, resp_ref_file)
    with open(raw_file, "r") as f:
        for line in f:
            item = list(json.loads(line.strip()).values())[0]
            history, search_text = [], ""
            for turn in item["dialog_history"]:
                action = turn["action"]
                if action == "Wizard => SearchAgent":
                    search_text = turn["text"]
                elif action == "Apprentice => Wizard":
                    history.append(turn["text"])
                elif action == "Wizard => Apprentice":
                    if not history:
                        history.append(turn["text"])
                        continue
                    contents, selects = turn["context"]["contents"], turn["context"]["selected_contents"]
                    no_knowledge, selects = selects[0][0], selects[1:]
                    assert len(selects) == len(contents)
                    knowledge = ""
                    if not no_knowledge:
                        for content, select in zip(contents, selects):
                            assert len(content["content"]) == len(select)
                            knowledge = next((c for c, s in zip(content["content"], select) if s), "")
                            if knowledge:
                                break
                    topic = search_text if knowledge else "no_topic"
                    response = strip(turn["text"])
                    if topic != "no_topic":
                        out.emit(strip(topic), strip(" [SEP] ".join(history)), strip(knowledge), response)
                    history.append(response)
                else:
                    assert action == "SearchAgent => Wizard", "Please check whether you have used the correct data!"
    out.close()


def get_database(test_datapath, train_datapath, data_type):
    """Training examples grouped by topic (topics seen in the test set) + the flat list used for unseen topics."""
    assert data_type in ["wow_seen", "wow_unseen", "woi"], "Please input a correct data type!!"
    with open(test_datapath, "r") as f:
        test_topics = {line.strip().split("\t")[0] for line in f}
    by_topic, dialogs_by_topic, examples = {}, {}, []
    with open(train_datapath, "r") as f:
        for line in f:
            topic, context, knowledge, _ = line.strip().split("\t")[:4]
            turns = context.split(" [SEP] ")[-3:]
            if knowledge == NO_PASSAGE:
                continue
            if data_type != "wow_seen" and ("(" in knowledge or ")" in knowledge or topic not in knowledge):
                continue
            instance = "( " + turns[-1] + " ) " + topic + " => " + knowledge
            dialog = ("( " + topic + " ) " if data_type != "wow_seen" else "") + " ".join(turns)
            if topic in test_topics:
                by_topic.setdefault(topic, []).append(instance)
                dialogs_by_topic.setdefault(topic, []).append(dialog)
            elif len(knowledge.split()) > 20 or knowledge.lower().startswith(("it", "this")):
                continue
            examples.append((topic, dialog, instance))
    return by_topic, dialogs_by_topic, examples


_emb_cache = {}


def _embed(texts, tokenizer, encoder, device):
    with torch.no_grad():
        return torch.cat([encoder(input_ids=torch.tensor([tokenizer.encode(t)], device=device)).pooler_output
                          for t in texts], dim=0)


def select_prompts_based_on_similarity(query, dialog_list, prompt_list, topic, tokenizer, encoder, topk):
    """The ``topk`` training dialogues closest to the query, most similar last."""
    device = next(encoder.parameters()).device
    q = _embed([query], tokenizer, encoder, device)[0]
    if topic not in _emb_cache:
        _emb_cache[topic] = _embed(dialog_list, tokenizer, encoder, device).cpu()
    sims = _emb_cache[topic].to(device).matmul(q)
    idx = torch.topk(sims, k=topk).indices.tolist()[::-1]
    return [prompt_list[i] for i in idx]


def prompt_selection_for_knowledge_generation(test_datapath, train_datapath, model_path, output_prompt_path,
                                              data_type):
    from transformers import DPRQuestionEncoderTokenizer
    by_topic, dialogs_by_topic, examples = get_database(test_datapath, train_datapath, data_type)
    tokenizer = DPRQuestionEncoderTokenizer.from_pretrained("facebook/dpr-question_encoder-single-nq-base")
    device = "cuda" if torch.cuda.is_available() else "cpu"
    encoder = torch.load(model_path, weights_only=False).to(device)
    all_emb = _embed([e[1] for e in examples], tokenizer, encoder, device)
    rows = []
    with open(test_datapath, "r") as f:
        for line in f:
            splits = line.strip().split("\t")
            topic, turns = splits[0], splits[1].split(" [SEP] ")[-3:]
            query = ("( " + topic + " ) " if data_type != "seen" else "") + " ".join(turns)
            if topic in by_topic:
                k = min(len(by_topic[topic]), 10)
                chosen = select_prompts_based_on_similarity(query, dialogs_by_topic[topic], by_topic[topic], topic,
                                                            tokenizer, encoder, topk=k)
            else:     # unseen topic: 10 examples from distinct topics
                sims = all_emb.matmul(_embed([query], tokenizer, encoder, device)[0])
                seen, chosen = set(), []
                for i in torch.sort(sims).indices.tolist():
                    if examples[i][0] not in seen:
                        seen.add(examples[i][0])
                        chosen.append(examples[i][2])
                        if len(chosen) == 10:
                            break
                chosen = chosen[::-1]
            rows.append({topic + " " + turns[-1]: chosen})
    with open(output_prompt_path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


def _overlap_tokens(response_tokens, knowledge_vocab, min_run=10):
    """Number of response tokens inside runs (>= ``min_run`` long) of tokens that also occur in the knowledge."""
    total = run = 0
    for tok in response_tokens + [None]:
        if tok is not None and tok in knowledge_vocab:
            run += 1
        else:
            total += run if run >= min_run else 0
            run = 0
    return total


def prompt_selection_for_response_generation(input_path, output_path, seed):
    """20 random training samples whose response copies most (60-90%) of its knowledge sentence."""
    np.random.seed(seed)
    prompts = []
    with open(input_path, "r") as f:
        for line in f:
            topic, context, knowledge, response = line.strip().split("\t")[:4]
            if knowledge == NO_PASSAGE:
                continue
            k_tok, r_tok = word_tokenize(knowledge), word_tokenize(response)
            overlap = _overlap_tokens(r_tok, set(k_tok))
            if not (0.6 * len(r_tok) <= overlap <= 0.9 * len(r_tok)) or overlap < 0.8 * len(k_tok):
                continue
            last_turn = " ".join(word_tokenize(context.split(" [SEP] ")[-1]))
            prompts.append("Topic: " + topic + ". User says: " + last_turn + " We know that: " + " ".join(k_tok)
                           + " System replies: " + " ".join(r_tok))
    np.random.shuffle(prompts)
    with open(output_path, "w") as f:
        for p in prompts[:20]:
            f.write(p + "\n")


def prepare_input_for_response_generation(test_file, knwl_gen_file, processed_file):
    """Swap the golden knowledge of every test sample for the generated one."""
    with open(knwl_gen_file, "r") as f:
        generated = [k.strip().replace("<|endoftext|>", "") for k in f]
    with open(test_file, "r") as fr, open(processed_file, "w") as fw:
        for i, line in enumerate(fr):
            s = line.strip().split("\t")
            fw.write("\t".join([s[0], s[1], generated[i], s[3]]) + "\n")


if __name__ == "__main__":
    a = get_args()
    if a.func == "process_wow_dataset":
        process_wow_dataset(a.raw_file, a.processed_file, a.knwl_ref_file, a.resp_ref_file)
    elif a.func == "process_woi_dataset":
        process_woi_dataset(a.raw_file, a.processed_file, a.knwl_ref_file, a.resp_ref_file)
    elif a.func == "get_knwl_gen_prompts":
        prompt_selection_for_knowledge_generation(a.test_file, a.train_file, a.model_file, a.processed_file, a.data_type)
    elif a.func == "get_resp_gen_prompts":
        prompt_selection_for_response_generation(a.train_file, a.processed_file, a.seed)
    elif a.func == "prepare_input":
        prepare_input_for_response_generation(a.test_file, a.knwl_gen_file, a.processed_file)
