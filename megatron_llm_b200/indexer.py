"""Evidence index builder: one pass of the context encoder over the Wikipedia evidence, embeddings sharded per
data-parallel rank and merged by rank 0.  Parity: megatron/indexer.py."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .checkpointing import load_biencoder_checkpoint
from .data.biencoder_dataset_utils import get_one_epoch_dataloader
from .data.orqa_wiki_dataset import get_open_retrieval_batch, get_open_retrieval_wiki_dataset
from .data.realm_index import OpenRetreivalDataStore, detach
from .models.biencoder_model import get_model_provider
from .parallel import state as ps


class IndexBuilder:
    def __init__(self, args):
        self.args = args
        self.biencoder_shared_query_context_model = args.biencoder_shared_query_context_model
        assert not (args.load and args.ict_load)
        self.log_interval, self.batch_size = args.indexer_log_interval, args.indexer_batch_size
        self.load_attributes(args)
        self.is_main_builder = ps.get_data_parallel_rank() == 0
        self.num_total_builders = ps.get_data_parallel_world_size()
        self.iteration = self.total_processed = 0

    def load_attributes(self, args):
        """Model (context tower only unless the towers are shared), one-epoch dataloader, empty store."""
        from .training import get_model
        only_context = not self.biencoder_shared_query_context_model
        provider = get_model_provider(only_context_model=only_context,
                                      biencoder_shared_query_context_model=self.biencoder_shared_query_context_model)
        model = get_model(provider, args=args)
        self.model = load_biencoder_checkpoint(model, only_context_model=only_context)
        assert len(self.model) == 1
        self.model[0].eval()
        self.dataset = get_open_retrieval_wiki_dataset()
        self.dataloader = iter(get_one_epoch_dataloader(self.dataset, self.batch_size))
        self.evidence_embedder_obj = OpenRetreivalDataStore(load_from_path=False)

    def track_and_report_progress(self, batch_size):
        self.iteration += 1
        self.total_processed += batch_size * self.num_total_builders
        if self.is_main_builder and self.iteration % self.log_interval == 0:
            print("Batch {:10d} | Total {:10d}".format(self.iteration, self.total_processed), flush=True)

    @torch.no_grad()
    def build_and_save_index(self):
        model = self.model[0]
        while not hasattr(model, "embed_text"):
            model = model.module
        while True:
            try:
                row_id, tokens, mask, types, _ = get_open_retrieval_batch(self.dataloader)
            except (StopIteration, IndexError):
                break
            assert mask.dtype == torch.bool
            logits = model.embed_text(model.context_model, tokens, mask, types)
            self.evidence_embedder_obj.add_block_data(detach(row_id), detach(logits.float()))
            self.track_and_report_progress(batch_size=len(row_id))
        self.evidence_embedder_obj.save_shard()
        dist.barrier()
        del self.model
        if self.is_main_builder:
            self.evidence_embedder_obj.merge_shards_and_save()
            assert len(self.evidence_embedder_obj.embed_data) == len(self.dataset)
        self.evidence_embedder_obj.clear()
        dist.barrier()
