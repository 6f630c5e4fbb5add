"""Merge every indexed dataset (``*.bin`` + ``*.idx`` pairs) found in a directory into one.

Parity: tools/merge_datasets.py."""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), os.path.pardir)))
from megatron_llm_b200.data import indexed_dataset  # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument_group(title="input data").add_argument(
        "--input", type=str, required=True, help="Path to directory containing all document files to merge")
    parser.add_argument_group(title="output data").add_argument(
        "--output_prefix", type=str, required=True, help="Path to binary output file without suffix")
    args = parser.parse_args(argv)
    assert os.path.isdir(args.input), f"ERROR: {args.input} is not a directory or does not exist"
    assert os.path.isdir(os.path.dirname(os.path.abspath(args.output_prefix))), \
        f"ERROR: {os.path.dirname(args.output_prefix)} is not a directory or does not exist"

    prefixes = set()
    for name in sorted(os.listdir(args.input)):
        stem, ext = os.path.splitext(name)
        if ext not in (".bin", ".idx") or stem in prefixes:
            continue
        for other in (".bin", ".idx"):
            assert os.path.isfile(os.path.join(args.input, stem + other)), \
                f"ERROR: {stem + other} does not exist in {args.input}"
        prefixes.add(stem)

    builder = None
    for stem in sorted(prefixes):
        path = os.path.join(args.input, stem)
        if builder is None:
            ds = indexed_dataset.make_dataset(path, "infer")
            cls = indexed_dataset.MMapIndexedDatasetBuilder if isinstance(ds, indexed_dataset.MMapIndexedDataset) \
                else indexed_dataset.IndexedDatasetBuilder
            dtype = ds._index.dtype if isinstance(ds, indexed_dataset.MMapIndexedDataset) else ds.dtype
            builder = cls(indexed_dataset.data_file_path(args.output_prefix), dtype=dtype)
            del ds
        builder.merge_file_(path)
    assert builder is not None, "no datasets found"
    builder.finalize(indexed_dataset.index_file_path(args.output_prefix))


if __name__ == "__main__":
    main()
