"""Retriever evaluation: build the evidence index with the context encoder, then score NQ questions
(parity: tasks/orqa/evaluate_orqa.py)."""
from megatron_llm_b200 import get_args, print_rank_0
from megatron_llm_b200.indexer import IndexBuilder
from tasks.orqa.evaluate_utils import ORQAEvaluator


def main():
    args = get_args()
    print_rank_0("Starting index builder!")
    IndexBuilder(args).build_and_save_index()
    print_rank_0("Build and save indices: done!")
    print_rank_0("Starting evaluations!")
    evaluator = ORQAEvaluator()
    if args.qa_data_dev is not None:
        evaluator.evaluate(args.qa_data_dev, "DEV")
    if args.qa_data_test is not None:
        evaluator.evaluate(args.qa_data_test, "TEST")
