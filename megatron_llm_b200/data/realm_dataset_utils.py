"""REALM dataset utilities (parity: megatron/data/realm_dataset_utils.py) -- same code as the bi-encoder utilities."""
from .biencoder_dataset_utils import (BlockSampleData, BlockSamplesMapping, get_block_samples_mapping,  # noqa: F401
                                      get_ict_batch, get_one_epoch_dataloader, join_str_list)
