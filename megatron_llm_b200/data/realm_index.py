"""Evidence-embedding store and exact maximum-inner-product search.

Parity: megatron/data/realm_index.py.  ``OpenRetreivalDataStore`` keeps the per-rank shard / merge protocol and pickle
format.  The reference delegates the search to FAISS ``IndexFlatIP``; exact MIPS is one GEMM + top-k, so
``FaissMIPSIndex`` here runs it on the GPU with torch (chunked over the evidence so any corpus fits) and needs no
extra dependency."""
from __future__ import annotations

import os
import pickle
import shutil

import numpy as np
import torch

from ..parallel import state as ps


def detach(tensor):
    return tensor.detach().cpu().numpy()


def _is_main():
    return (not ps.model_parallel_is_initialized()) or ps.get_data_parallel_rank() == 0


class OpenRetreivalDataStore:
    """row id -> fp16 embedding, serialisable; every rank saves a shard, rank 0 merges them."""

    def __init__(self, embedding_path=None, load_from_path=True, rank=None):
        self.embed_data = {}
        if embedding_path is None:
            from ..global_vars import get_args
            args = get_args()
            embedding_path, rank = args.embedding_path, args.rank
        self.embedding_path, self.rank = embedding_path, rank
        if load_from_path:
            self.load_from_file()
        self.temp_dir_name = os.path.splitext(self.embedding_path)[0] + "_tmp"

    def state(self):
        return {"embed_data": self.embed_data}

    def clear(self):
        self.embed_data = {}

    def load_from_file(self):
        if _is_main():
            print("\n> Unpickling BlockData", flush=True)
        with open(self.embedding_path, "rb") as f:
            self.embed_data = pickle.load(f)["embed_data"]
        if _is_main():
            print(">> Finished unpickling BlockData\n", flush=True)

    def add_block_data(self, row_id, block_embeds, allow_overwrite=False):
        for idx, embed in zip(row_id, block_embeds):
            idx = int(idx)
            if not allow_overwrite and idx in self.embed_data:
                raise ValueError("Unexpectedly tried to overwrite block data")
            self.embed_data[idx] = np.float16(embed)

    def save_shard(self):
        os.makedirs(self.temp_dir_name, exist_ok=True)
        with open(f"{self.temp_dir_name}/{self.rank}.pkl", "wb") as f:
            pickle.dump(self.state(), f)

    def merge_shards_and_save(self):
        names = os.listdir(self.temp_dir_name)
        seen_own = False
        for fname in names:
            if int(os.path.splitext(fname)[0]) == self.rank:
                seen_own = True
                continue
            with open(f"{self.temp_dir_name}/{fname}", "rb") as f:
                shard = pickle.load(f)["embed_data"]
            before = len(self.embed_data)
            self.embed_data.update(shard)
            assert len(self.embed_data) == before + len(shard)
        assert seen_own
        with open(self.embedding_path, "wb") as f:
            pickle.dump(self.state(), f)
        shutil.rmtree(self.temp_dir_name, ignore_errors=True)
        print(f"Finished merging {len(names)} shards for a total of {len(self.embed_data)} embeds", flush=True)


class FaissMIPSIndex:
    """Exact inner-product top-k over the evidence embeddings (same interface as the reference's FAISS wrapper)."""

    def __init__(self, embed_size, embed_data=None, use_gpu=False, chunk=1 << 20):
        self.embed_size, self.embed_data, self.chunk = embed_size, embed_data, chunk
        self.device = torch.device("cuda", torch.cuda.current_device()) if use_gpu and torch.cuda.is_available() \
            else torch.device("cpu")
        self.use_gpu = use_gpu
        self._set_mips_index()

    def _set_mips_index(self):
        if _is_main():
            print("\n> Building index", flush=True)
        dtype = torch.float16 if self.device.type == "cuda" else torch.float32
        self.embeds = torch.empty(0, self.embed_size, dtype=dtype, device=self.device)
        self.ids = torch.empty(0, dtype=torch.long, device=self.device)
        if self.embed_data is not None:
            self.add_embed_data(self.embed_data)

    def reset_index(self):
        if self.embed_data is not None:
            self.embed_data = OpenRetreivalDataStore(self.embed_data.embedding_path)
        self._set_mips_index()

    def update_index(self):
        if self.embed_data is not None:
            self.embed_data.load_from_file()
        self._set_mips_index()

    def add_embed_data(self, all_embed_data):
        ids, embeds = zip(*all_embed_data.embed_data.items())
        arr = torch.from_numpy(np.asarray(embeds, dtype=np.float32)).to(self.device, self.embeds.dtype)
        self.embeds = torch.cat([self.embeds, arr])
        self.ids = torch.cat([self.ids, torch.tensor(ids, dtype=torch.long, device=self.device)])
        all_embed_data.clear()
        if _is_main():
            print(">>> Finished adding block data to index", flush=True)

    def search_mips_index(self, query_embeds, top_k, reconstruct=True):
        """reconstruct=True -> [queries, k, dim] embeddings; else (scores [queries, k], ids [queries, k])."""
        q = torch.as_tensor(np.float32(detach(query_embeds)) if torch.is_tensor(query_embeds) else query_embeds)
        q = q.to(self.device, self.embeds.dtype)
        best_s = best_i = None
        for lo in range(0, self.embeds.size(0), self.chunk):
            s = (q @ self.embeds[lo:lo + self.chunk].t()).float()
            k = min(top_k, s.size(1))
            cs, ci = torch.topk(s, k, dim=1)
            ci = ci + lo
            if best_s is None:
                best_s, best_i = cs, ci
            else:
                cat_s, cat_i = torch.cat([best_s, cs], 1), torch.cat([best_i, ci], 1)
                best_s, sel = torch.topk(cat_s, min(top_k, cat_s.size(1)), dim=1)
                best_i = torch.gather(cat_i, 1, sel)
        if reconstruct:
            return self.embeds[best_i].float().cpu().numpy()
        return best_s.cpu().numpy(), self.ids[best_i].cpu().numpy()
