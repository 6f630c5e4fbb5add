"""MSDP-PROMPT: few-shot prompting of a pretrained LM for knowledge / response generation
(parity: tasks/msdp/prompt.py).  Either through the in-process model or the REST server (``--api_prompt``)."""
import json
import re
import urllib.request

import torch

from megatron_llm_b200 import get_args, print_rank_0
from megatron_llm_b200.checkpointing import load_checkpoint
from megatron_llm_b200.parallel import state as mpu
from megatron_llm_b200.text_generation import generate_and_post_process
from megatron_llm_b200.training import get_model


def word_tokenize(text):
    """NLTK's tokenizer when available, a punctuation-splitting regex otherwise."""
    try:
        from nltk import word_tokenize as wt
        return wt(text)
    except Exception:
        return re.findall(r"\w+(?:'\w+)?|[^\w\s]", text)


def call_model_api(inputs, tokens_to_generate):
    args = get_args()
    body = json.dumps({"prompts": [inputs], "tokens_to_generate": tokens_to_generate, "top_k": 1}).encode()
    req = urllib.request.Request(args.megatron_api_url, data=body, method="PUT",
                                 headers={"Content-Type": "application/json; charset=UTF-8"})
    with urllib.request.urlopen(req) as resp:
        text = json.loads(resp.read())["text"][0]
    return text[len(inputs):].split("\n")[0].strip()


def read_prompts(prompt_path, prompt_type, n_example):
    """knowledge: json lines {"topic last_turn": [examples]} -> dict of prompt strings; response: first n lines."""
    join = lambda xs: "".join(x.strip() + " \n" for x in xs)       # noqa: E731
    with open(prompt_path, "r") as f:
        if prompt_type != "knowledge":
            return join(f.readlines()[:n_example])
        prompts = {}
        for line in f:
            d = json.loads(line.strip())
            key = next(iter(d))
            prompts.setdefault(key, join(d[key]))
        return prompts


def build_input(sample, prompt_type, prompts):
    """One tab-separated test sample -> the full prompt text."""
    splits = sample.strip().split("\t")
    topic, last_turn = splits[0], splits[1].split(" [SEP] ")[-1]
    if prompt_type == "knowledge":
        return prompts[topic + " " + last_turn] + "( " + last_turn + " ) " + topic + " =>"
    last_turn = " ".join(word_tokenize(last_turn)).strip()
    knowledge = " ".join(word_tokenize(splits[2])).strip()
    return (prompts + "Topic: " + topic + ". " + "User says: " + last_turn + " " + "We know that: " + knowledge + " "
            + "System replies:")


def generate_samples_by_calling_api():
    args = get_args()
    assert args.prompt_type in ["knowledge", "response"], "Please input a correct prompt type!"
    prompts = read_prompts(args.prompt_file, args.prompt_type, args.num_prompt_examples)
    with open(args.sample_input_file, "r") as fin, open(args.sample_output_file, "w") as fout:
        for sample in fin:
            fout.write(call_model_api(build_input(sample, args.prompt_type, prompts), args.out_seq_length) + "\n")


def model_provider(pre_process=True, post_process=True):
    import finetune
    print_rank_0("building model for prompting ...")
    model = finetune.model_provider(pre_process, post_process)
    # sampling needs the full-vocabulary logits on every TP rank (the reference builds this model with
    # parallel_output=True, tasks/msdp/prompt.py:21-26, which only works for TP = 1)
    from megatron_llm_b200.utils import unwrap_model
    unwrap_model(model).parallel_output = False
    return model


def generate_samples_by_prompting_input_from_file(model):
    args = get_args()
    assert args.sample_input_file is not None, "sample input file is not provided."
    assert args.prompt_type in ["knowledge", "response"], "Please input a correct prompt type!"
    writer = mpu.is_pipeline_first_stage() and mpu.get_tensor_model_parallel_rank() == 0
    with open(args.sample_input_file, "r") as f:
        samples = f.readlines()
    prompts = read_prompts(args.prompt_file, args.prompt_type, args.num_prompt_examples)
    fout = None
    if writer:
        out_path = args.sample_output_file or args.sample_input_file + ".out"
        if args.sample_output_file is None:
            print("`sample_output_file` not specified, setting it to {}".format(out_path))
        fout = open(out_path, "w")
    model.eval()
    with torch.no_grad():
        for pos, sample in enumerate(samples, 1):
            raw_text = build_input(sample, args.prompt_type, prompts) if writer else "EMPTY TEXT"
            if pos % 100 == 0:
                print_rank_0("input_pos: %d" % pos)
            outputs = generate_and_post_process(model=model, prompts=[raw_text],
                                                tokens_to_generate=args.out_seq_length, top_k_sampling=1)
            if writer:
                fout.write(outputs[0][0][len(raw_text):].split("\n")[0].strip() + "\n")
    if fout:
        fout.close()


def main():
    args = get_args()
    if args.api_prompt:
        return generate_samples_by_calling_api()
    if args.num_layers_per_virtual_pipeline_stage is not None:
        print("Interleaved pipeline schedule is not yet supported for text generation.")
        return
    model = get_model(model_provider, wrap_with_ddp=False, args=args)
    if args.load is not None:
        load_checkpoint(model, None, None)
    assert len(model) == 1, "Above condition should have caught this"
    generate_samples_by_prompting_input_from_file(model[0])
