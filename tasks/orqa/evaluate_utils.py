"""Open-retrieval QA evaluator: encode the questions, exact MIPS over the evidence embeddings on the node's first
rank, broadcast the top-k, count answer hits (parity: tasks/orqa/evaluate_utils.py)."""
import torch
import torch.distributed as dist

from megatron_llm_b200 import get_args, print_rank_0
from megatron_llm_b200.checkpointing import load_biencoder_checkpoint
from megatron_llm_b200.data.orqa_wiki_dataset import get_open_retrieval_wiki_dataset
from megatron_llm_b200.data.realm_index import FaissMIPSIndex, OpenRetreivalDataStore
from megatron_llm_b200.models.biencoder_model import get_model_provider
from megatron_llm_b200.models.enums import ModelType
from megatron_llm_b200.training import get_model
from megatron_llm_b200.utils.device import current_device
from tasks.orqa.unsupervised.nq import get_nq_dataset, get_one_epoch_nq_dataloader, process_nq_batch
from tasks.orqa.unsupervised.qa_utils import calculate_matches


class ORQAEvaluator:
    def __init__(self):
        args = get_args()
        self.embedding_size = args.hidden_size if args.biencoder_projection_dim == 0 else args.biencoder_projection_dim
        self.faiss_use_gpu = args.faiss_use_gpu
        self.evidence_embedder_obj = self.mips_index = self.eval_dataset = None
        self.evidence_dataset = get_open_retrieval_wiki_dataset()
        only_query = not args.biencoder_shared_query_context_model
        provider = get_model_provider(only_query_model=only_query,
                                      biencoder_shared_query_context_model=args.biencoder_shared_query_context_model,
                                      model_type=ModelType.encoder_or_decoder)
        model = get_model(provider, ModelType.encoder_or_decoder, True, args)
        self.model = load_biencoder_checkpoint(model, only_query_model=only_query)
        assert len(self.model) == 1
        self.model[0].eval()
        self.faiss_wrapper()

    def faiss_wrapper(self):
        """The first rank of every node owns the (exact, GEMM + top-k) index."""
        args = get_args()
        if (args.local_rank or 0) == 0:
            self.evidence_embedder_obj = OpenRetreivalDataStore(load_from_path=True)
            self.mips_index = FaissMIPSIndex(embed_size=self.embedding_size, embed_data=self.evidence_embedder_obj,
                                             use_gpu=self.faiss_use_gpu)
        dist.barrier()

    @torch.no_grad()
    def generate_query_vectors(self, qa_data, split):
        self.eval_dataset = get_nq_dataset(qa_data, split)
        model = self.model[0]
        while not hasattr(model, "embed_text"):
            model = model.module
        vectors, references = [], []
        for batch in get_one_epoch_nq_dataloader(self.eval_dataset):
            tokens, mask, types, _, reference = process_nq_batch(batch)
            vectors.append(model.embed_text(model.query_model, tokens, mask, types))
            references.extend(reference)
        query_tensor = torch.cat(vectors, dim=0)
        print_rank_0("Total encoded queries tensor {}".format(query_tensor.size()))
        assert query_tensor.size(0) == len(self.eval_dataset)
        return query_tensor, references

    def evaluate(self, qa_data, split):
        args = get_args()
        query_tensor, reference_list = self.generate_query_vectors(qa_data, split)
        local_rank = args.local_rank or 0
        rank, world = dist.get_rank(), dist.get_world_size()
        per_node = max(1, min(world, torch.cuda.device_count() if torch.cuda.is_available() else world))
        node_id, group, first = rank // per_node, None, 0
        for node in range(world // per_node):          # every rank must create every group
            ranks = list(range(node * per_node, (node + 1) * per_node))
            g = dist.new_group(ranks=ranks)
            if node == node_id:
                group, first = g, ranks[0]
        nq, k = query_tensor.size(0), args.faiss_topk_retrievals
        gathered = torch.empty((per_node * nq, query_tensor.size(1)), dtype=query_tensor.dtype,
                               device=query_tensor.device)
        dist.all_gather_into_tensor(gathered, query_tensor.contiguous(), group=group)
        dev = current_device()
        if local_rank == 0 and self.mips_index is not None:
            distance, topk = self.mips_index.search_mips_index(gathered, top_k=k, reconstruct=False)
            distance, topk = torch.from_numpy(distance).float().to(dev), torch.from_numpy(topk).long().to(dev)
            k = distance.size(1)
        else:
            distance = torch.empty(per_node * nq, k, dtype=torch.float32, device=dev)
            topk = torch.empty(per_node * nq, k, dtype=torch.int64, device=dev)
        dist.broadcast(distance, src=first, group=group)
        dist.broadcast(topk, src=first, group=group)
        distance, topk = distance.split(nq, dim=0)[local_rank], topk.split(nq, dim=0)[local_rank]
        top_ids_and_scores = [(ids.tolist(), d.tolist()) for d, ids in zip(distance, topk)]
        stats = calculate_matches(self.evidence_dataset.id2text, reference_list, top_ids_and_scores,
                                  workers_num=args.num_workers, match_type=args.faiss_match)
        print_rank_0("{} SET RESULTS".format(split))
        print_rank_0("topk-{} documents hits {}".format(args.faiss_topk_retrievals, stats.top_k_hits))
        acc = [v / len(top_ids_and_scores) for v in stats.top_k_hits]
        print_rank_0("top-k documents hits accuracy {}".format(acc))
        for i in args.retriever_report_topk_accuracies:
            if i - 1 < len(acc):
                print_rank_0("top-{}: {:.2f}".format(i, acc[i - 1] * 100))
        return acc
