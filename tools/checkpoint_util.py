"""Reshard a Megatron checkpoint to a different tensor/pipeline-parallel layout.

Parity: tools/checkpoint_util.py -- same CLI (``--model_type --loader --saver --load_dir --save_dir
--target_tensor_parallel_size --target_pipeline_parallel_size --true_vocab_size --bf16 ...``) and the same
loader -> queue -> saver plugin protocol:

    metadata namespace, then dict messages (each with a "name"):
      "embeddings"            {"word embeddings", ["position embeddings"], ["tokentype embeddings"]}
      "lm_head"               {"lm_head"}                       (only when embeddings are untied)
      "transformer layer N"   {"input layernorm weight/bias", ["mlp layernorm weight/bias"], "qkv weight/bias",
                               "dense weight/bias", "post layernorm weight/bias", "mlp l0 weight/bias",
                               "mlp l1 weight/bias"}             (full, unsharded tensors; GLU l0 = [up; gate])
      "final layernorm"       {"weight", ["bias"]}
      BERT only: "pooler", "lm head", "binary head"
    then the string "done" ("exit" = the loader failed).

Plugins are modules ``checkpoint_loader_<name>`` / ``checkpoint_saver_<name>`` exposing ``add_arguments`` and
``load_checkpoint(queue, args)`` / ``save_checkpoint(queue, args)``.  The loader runs in a thread feeding a bounded
queue (``--max_queue_size``) so reading and writing overlap without spawning a second process."""
from __future__ import annotations

import argparse
import importlib
import os
import queue as queue_mod
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))


def load_plugin(plugin_type, name):
    for module_name in (f"checkpoint_{plugin_type}_{name}", name):
        try:
            plugin = importlib.import_module(module_name)
        except ModuleNotFoundError:
            continue
        if not hasattr(plugin, "add_arguments"):
            sys.exit(f"{module_name} module is not a plugin. Exiting.")
        print(f"Loaded {module_name} as the {plugin_type}.")
        return plugin
    sys.exit(f"Unable to load {plugin_type} plugin {name}. Exiting.")


def main(argv=None):
    parser = argparse.ArgumentParser(description="Megatron checkpoint utility: change the TP/PP layout",
                                     allow_abbrev=False, conflict_handler="resolve")
    parser.add_argument("--model_type", type=str, required=True,
                        choices=["GPT", "BERT", "falcon", "llama", "llama2", "codellama", "mistral"],
                        help="Type of the model")
    parser.add_argument("--loader", type=str, default="megatron", help="Module name to load checkpoint")
    parser.add_argument("--saver", type=str, default="megatron", help="Module name to save checkpoint")
    parser.add_argument("--load_dir", type=str, required=True, help="Directory to load model checkpoint from")
    parser.add_argument("--save_dir", type=str, required=True, help="Directory to save model checkpoint to")
    parser.add_argument("--max_queue_size", type=int, default=50, help="Maximum number of tensors in the queue")
    parser.add_argument("--no_checking", action="store_false", dest="checking",
                        help="Do not perform checking on the name and ordering of weights")
    parser.add_argument("--bf16", action="store_true", help="force bfloat16 weights")
    parser.add_argument("--load_iters", type=int, default=None, help="iteration to load (default: latest)")
    known_args, _ = parser.parse_known_args(argv)
    loader = load_plugin("loader", known_args.loader)
    saver = load_plugin("saver", known_args.saver)
    loader.add_arguments(parser)
    saver.add_arguments(parser)
    args = parser.parse_args(argv)

    q = queue_mod.Queue(maxsize=args.max_queue_size)
    failure = []

    def produce():
        try:
            loader.load_checkpoint(q, args)
        except BaseException as e:  # the loader already put "exit"
            failure.append(e)

    print("Starting saver...")
    t = threading.Thread(target=produce, daemon=True)
    t.start()
    saver.save_checkpoint(q, args)
    t.join()
    if failure:
        raise failure[0]


if __name__ == "__main__":
    main()
