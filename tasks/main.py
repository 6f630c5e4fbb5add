"""Downstream-task dispatcher (parity: tasks/main.py:74-96): ``python tasks/main.py --task <NAME> ...``.

Tasks: RACE, MNLI, QQP (fine-tune + accuracy), LAMBADA, WIKITEXT103 (zero-shot GPT), ICT-ZEROSHOT-NQ, RETRIEVER-EVAL
(retriever evaluation), RET-FINETUNE-NQ (supervised retriever fine-tuning)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))

from megatron_llm_b200 import get_args  # noqa: E402
from megatron_llm_b200.initialize import initialize_megatron  # noqa: E402


def get_tasks_args(parser):
    group = parser.add_argument_group(title="tasks")
    group.add_argument("--task", type=str, required=True, help="Task name.")
    group.add_argument("--epochs", type=int, default=None,
                       help="Number of finetunning epochs. Zero results in evaluation only.")
    group.add_argument("--pretrained_checkpoint", type=str, default=None,
                       help="Pretrained checkpoint used for finetunning.")
    group.add_argument("--keep_last", action="store_true",
                       help="Keep the last batch (maybe incomplete) in the data loader")
    group.add_argument("--train_data", nargs="+", default=None,
                       help="Whitespace separated paths or corpora names for training.")
    group.add_argument("--valid_data", nargs="*", default=None, help="path(s) to the validation data.")
    group.add_argument("--overlapping_eval", type=int, default=32, help="Sliding window for overlapping evaluation.")
    group.add_argument("--strict_lambada", action="store_true", help="Use more difficult formulation of lambada.")
    group.add_argument("--qa_data_dev", type=str, default=None, help="Path to the QA dataset dev file.")
    group.add_argument("--qa_data_test", type=str, default=None, help="Path to the QA dataset test file.")
    group.add_argument("--faiss_use_gpu", action="store_true", help="Run the MIPS index on the GPU")
    group.add_argument("--faiss_match", type=str, default="string", choices=["regex", "string"],
                       help="Answer matching logic type")
    group.add_argument("--faiss_topk_retrievals", type=int, default=100,
                       help="Number of blocks to use as top-k during retrieval")
    group.add_argument("--eval_micro_batch_size", type=int, default=None,
                       help="Eval batch size per model instance (local batch size).")
    group.add_argument("--train_with_neg", action="store_true",
                       help="Whether to use negative examples during model training")
    group.add_argument("--train_hard_neg", type=int, default=0,
                       help="Number of hard negative exmaples to use during training")
    group.add_argument("--val_av_rank_hard_neg", type=int, default=30,
                       help="Av.rank validation: how many hard negatives to take from each question pool")
    group.add_argument("--val_av_rank_other_neg", type=int, default=30,
                       help="Av.rank validation: how many other negatives to take from each question pool")
    return parser


TASKS = {"RACE": "tasks.race.finetune", "MNLI": "tasks.glue.finetune", "QQP": "tasks.glue.finetune",
         "LAMBADA": "tasks.zeroshot_gpt.evaluate", "WIKITEXT103": "tasks.zeroshot_gpt.evaluate",
         "ICT-ZEROSHOT-NQ": "tasks.orqa.evaluate_orqa", "RETRIEVER-EVAL": "tasks.orqa.evaluate_orqa",
         "RET-FINETUNE-NQ": "tasks.orqa.supervised.finetune"}


def run(args_list=None):
    initialize_megatron(extra_args_provider=get_tasks_args, args_list=args_list)
    args = get_args()
    if args.num_layers_per_virtual_pipeline_stage is not None:
        print("Interleaved pipeline schedule is not yet supported for downstream tasks.")
        sys.exit()
    if args.task not in TASKS:
        raise NotImplementedError("Task {} is not implemented.".format(args.task))
    import importlib
    importlib.import_module(TASKS[args.task]).main()


if __name__ == "__main__":
    run()
