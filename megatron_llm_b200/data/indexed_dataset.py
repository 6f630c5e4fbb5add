"""Token-sequence datasets stored as ``<prefix>.bin`` (raw tokens) + ``<prefix>.idx`` (index).

Parity target: megatron/data/indexed_dataset.py.  Both on-disk formats of the reference are supported bit-for-bit:

* ``mmap``  (``MMIDIDX\\0\\0``): magic(9) | version u64=1 | dtype code u8 | n_seqs u64 | n_docs u64 |
  sizes int32[n_seqs] | pointers int64[n_seqs] (byte offsets) | doc_idx int64[n_docs]      (reference :341-545)
* ``lazy`` / ``cached`` (``TNTIDX\\0\\0``): magic(8) | version u64=1 | dtype code u64 | element size u64 | n u64 |
  s u64 | doc_count u64 | dim_offsets i64[n+1] | data_offsets i64[n+1] | sizes i64[s] | doc_idx i64[doc_count]

The class and function names, their order and the on-disk layout follow the reference (the formats are a compatibility
contract with existing preprocessed corpora); reads go through numpy memory maps.
"""
from __future__ import annotations

import os
import shutil
import struct
from functools import lru_cache
from itertools import accumulate

import numpy as np
import torch

from ..utils import print_rank_0

_DTYPES = {1: np.uint8, 2: np.int8, 3: np.int16, 4: np.int32, 5: np.int64, 6: float, 7: np.double, 8: np.uint16}
dtypes = _DTYPES  # reference name


def code(dtype):
    for k, v in _DTYPES.items():
        if v == dtype:
            return k
    raise ValueError(dtype)


def __best_fitting_dtype(vocab_size=None):
    return np.uint16 if (vocab_size is not None and vocab_size < 65500) else np.int32


best_fitting_dtype = __best_fitting_dtype


def get_available_dataset_impl():
    return ["lazy", "cached", "mmap"]


def index_file_path(prefix_path):
    return prefix_path + ".idx"


def data_file_path(prefix_path):
    return prefix_path + ".bin"


def infer_dataset_impl(path):
    if not IndexedDataset.exists(path):
        print(f"Dataset does not exist: {path}")
        print("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    with open(index_file_path(path), "rb") as f:
        magic = f.read(8)
    if magic == IndexedDataset._HDR_MAGIC:
        return "cached"
    if magic == MMapIndexedDataset.Index._HDR_MAGIC[:8]:
        return "mmap"
    return None


def make_builder(out_file, impl, vocab_size=None):
    if impl == "mmap":
        return MMapIndexedDatasetBuilder(out_file, dtype=best_fitting_dtype(vocab_size))
    return IndexedDatasetBuilder(out_file)


def make_dataset(path, impl: str, skip_warmup=False):
    if not IndexedDataset.exists(path):
        print(f"Dataset does not exist: {path}")
        print("Path should be a basename that both .idx and .bin can be appended to get full filenames.")
        return None
    if impl == "infer":
        impl = infer_dataset_impl(path)
    if impl == "lazy":
        return IndexedDataset(path)
    if impl == "cached":
        return IndexedCachedDataset(path)
    if impl == "mmap":
        return MMapIndexedDataset(path, skip_warmup)
    print(f"Unknown dataset implementation: {impl}")
    return None


def dataset_exists(path, impl):
    return MMapIndexedDataset.exists(path) if impl == "mmap" else IndexedDataset.exists(path)


def read_longs(f, n):
    a = np.empty(n, dtype=np.int64)
    f.readinto(a)
    return a


def write_longs(f, a):
    f.write(np.array(a, dtype=np.int64))


def create_doc_idx(sizes):
    """Documents end at zero-length sequences."""
    return [0] + [i + 1 for i, s in enumerate(sizes) if s == 0]


# ------------------------------------------------------------------------------------------------
# legacy "TNTIDX" format
# ------------------------------------------------------------------------------------------------

class IndexedDataset(torch.utils.data.Dataset):
    """Lazy loader: every ``__getitem__`` seeks and reads from the .bin file."""
    _HDR_MAGIC = b"TNTIDX\x00\x00"

    def __init__(self, path):
        super().__init__()
        self.path = path
        self.data_file = None
        self.read_index(path)

    def read_index(self, path):
        with open(index_file_path(path), "rb") as f:
            assert f.read(8) == self._HDR_MAGIC, \
                "Index file doesn't match expected format. Make sure that --dataset_impl is configured properly."
            assert struct.unpack("<Q", f.read(8)) == (1,)
            dcode, self.element_size = struct.unpack("<QQ", f.read(16))
            self.dtype = _DTYPES[dcode]
            self._len, self.s = struct.unpack("<QQ", f.read(16))
            (self.doc_count,) = struct.unpack("<Q", f.read(8))
            self.dim_offsets = read_longs(f, self._len + 1)
            self.data_offsets = read_longs(f, self._len + 1)
            self.sizes = read_longs(f, self.s)
            self.doc_idx = read_longs(f, self.doc_count)

    def read_data(self, path):
        self.data_file = open(data_file_path(path), "rb", buffering=0)

    def check_index(self, i):
        if i < 0 or i >= self._len:
            raise IndexError("index out of range")

    def __del__(self):
        if self.data_file:
            self.data_file.close()

    def _read(self, start_elem, count, shape=None):
        a = np.empty(count if shape is None else shape, dtype=self.dtype)
        self.data_file.seek(int(start_elem) * self.element_size)
        self.data_file.readinto(a)
        return a

    def __getitem__(self, idx):
        if not self.data_file:
            self.read_data(self.path)
        if isinstance(idx, (int, np.integer)):
            self.check_index(idx)
            shape = self.sizes[self.dim_offsets[idx]:self.dim_offsets[idx + 1]]
            return self._read(self.data_offsets[idx], None, tuple(int(s) for s in shape))
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("Slices into indexed_dataset must be contiguous")
            sizes = self.sizes[self.dim_offsets[start]:self.dim_offsets[stop]]
            flat = self._read(self.data_offsets[start], int(sum(sizes)))
            return np.split(flat, list(accumulate(sizes))[:-1])
        raise TypeError(type(idx))

    def __len__(self):
        return self._len

    def num_tokens(self, index):
        return self.sizes[index]

    def size(self, index):
        return self.sizes[index]

    @staticmethod
    def exists(path):
        return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))

    @property
    def supports_prefetch(self):
        return False


class IndexedCachedDataset(IndexedDataset):
    """Lazy dataset + an in-memory cache filled by ``prefetch(indices)``."""

    def __init__(self, path):
        super().__init__(path)
        self.cache = None
        self.cache_index = {}

    @property
    def supports_prefetch(self):
        return True

    def prefetch(self, indices):
        if all(i in self.cache_index for i in indices):
            return
        if not self.data_file:
            self.read_data(self.path)
        indices = sorted(set(indices))
        total = sum(int(self.data_offsets[i + 1] - self.data_offsets[i]) for i in indices)
        self.cache = np.empty(total, dtype=self.dtype)
        self.cache_index.clear()
        ptx = 0
        for i in indices:
            self.cache_index[i] = ptx
            n = int(self.data_offsets[i + 1] - self.data_offsets[i])
            self.data_file.seek(int(self.data_offsets[i]) * self.element_size)
            self.data_file.readinto(self.cache[ptx:ptx + n])
            ptx += n
        if self.data_file:
            self.data_file.close()
            self.data_file = None

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            self.check_index(idx)
            shape = tuple(int(s) for s in self.sizes[self.dim_offsets[idx]:self.dim_offsets[idx + 1]])
            a = np.empty(shape, dtype=self.dtype)
            ptx = self.cache_index[idx]
            np.copyto(a, self.cache[ptx:ptx + a.size].reshape(shape))
            return a
        if isinstance(idx, slice):
            return [self[i] for i in range(*idx.indices(len(self)))]
        raise TypeError(type(idx))


class IndexedDatasetBuilder:
    element_sizes = {np.uint8: 1, np.int8: 1, np.int16: 2, np.int32: 4, np.int64: 8, float: 4, np.double: 8}

    def __init__(self, out_file, dtype=np.int32):
        self.out_file = open(out_file, "wb")
        self.dtype = dtype
        self.data_offsets, self.dim_offsets, self.sizes = [0], [0], []
        self.element_size = self.element_sizes[self.dtype]
        self.doc_idx = [0]

    def add_item(self, tensor):
        nbytes = self.out_file.write(np.array(tensor.numpy(), dtype=self.dtype))
        self.data_offsets.append(self.data_offsets[-1] + nbytes // self.element_size)
        self.sizes.extend(tensor.size())
        self.dim_offsets.append(self.dim_offsets[-1] + len(tensor.size()))

    def end_document(self):
        self.doc_idx.append(len(self.sizes))

    def merge_file_(self, another_file):
        index = IndexedDataset(another_file)
        assert index.dtype == self.dtype
        doc_offset = len(self.sizes)
        base = self.data_offsets[-1]
        self.data_offsets.extend(base + int(o) for o in index.data_offsets[1:])
        self.sizes.extend(index.sizes)
        base = self.dim_offsets[-1]
        self.dim_offsets.extend(base + int(o) for o in index.dim_offsets[1:])
        self.doc_idx.extend(doc_offset + int(d) for d in index.doc_idx[1:])
        with open(data_file_path(another_file), "rb") as f:
            shutil.copyfileobj(f, self.out_file)

    def finalize(self, index_file):
        self.out_file.close()
        with open(index_file, "wb") as index:
            index.write(IndexedDataset._HDR_MAGIC)
            index.write(struct.pack("<Q", 1))
            index.write(struct.pack("<QQ", code(self.dtype), self.element_size))
            index.write(struct.pack("<QQ", len(self.data_offsets) - 1, len(self.sizes)))
            index.write(struct.pack("<Q", len(self.doc_idx)))
            write_longs(index, self.dim_offsets)
            write_longs(index, self.data_offsets)
            write_longs(index, self.sizes)
            write_longs(index, self.doc_idx)


# ------------------------------------------------------------------------------------------------
# mmap "MMIDIDX" format
# ------------------------------------------------------------------------------------------------

def _warmup_mmap_file(path):
    with open(path, "rb") as stream:
        while stream.read(100 * 1024 * 1024):
            pass


class MMapIndexedDataset(torch.utils.data.Dataset):
    class Index:
        _HDR_MAGIC = b"MMIDIDX\x00\x00"

        @classmethod
        def writer(cls, path, dtype):
            class _Writer:
                def __enter__(self):
                    self._file = open(path, "wb")
                    self._file.write(cls._HDR_MAGIC)
                    self._file.write(struct.pack("<Q", 1))
                    self._file.write(struct.pack("<B", code(dtype)))
                    return self

                @staticmethod
                def _get_pointers(sizes):
                    sizes64 = np.asarray(sizes, dtype=np.int64) * np.dtype(dtype).itemsize
                    ptrs = np.zeros(len(sizes64), dtype=np.int64)
                    if len(sizes64) > 1:
                        np.cumsum(sizes64[:-1], out=ptrs[1:])
                    return ptrs

                def write(self, sizes, doc_idx):
                    pointers = self._get_pointers(sizes)
                    self._file.write(struct.pack("<Q", len(sizes)))
                    self._file.write(struct.pack("<Q", len(doc_idx)))
                    self._file.write(np.array(sizes, dtype=np.int32).tobytes(order="C"))
                    self._file.write(pointers.tobytes(order="C"))
                    self._file.write(np.array(doc_idx, dtype=np.int64).tobytes(order="C"))

                def __exit__(self, exc_type, exc_val, exc_tb):
                    self._file.close()

            return _Writer()

        def __init__(self, path, skip_warmup=False):
            with open(path, "rb") as stream:
                assert stream.read(9) == self._HDR_MAGIC, \
                    "Index file doesn't match expected format. Make sure that --dataset_impl is configured properly."
                assert struct.unpack("<Q", stream.read(8)) == (1,)
                (dcode,) = struct.unpack("<B", stream.read(1))
                self._dtype = _DTYPES[dcode]
                self._dtype_size = np.dtype(self._dtype).itemsize
                (self._len,) = struct.unpack("<Q", stream.read(8))
                (self._doc_count,) = struct.unpack("<Q", stream.read(8))
                offset = stream.tell()
            if not skip_warmup:
                print_rank_0("    warming up index mmap file...")
                _warmup_mmap_file(path)
            self._bin_buffer_mmap = np.memmap(path, mode="r", order="C")
            self._bin_buffer = memoryview(self._bin_buffer_mmap)
            self._sizes = np.frombuffer(self._bin_buffer, dtype=np.int32, count=self._len, offset=offset)
            self._pointers = np.frombuffer(self._bin_buffer, dtype=np.int64, count=self._len,
                                           offset=offset + self._sizes.nbytes)
            self._doc_idx = np.frombuffer(self._bin_buffer, dtype=np.int64, count=self._doc_count,
                                          offset=offset + self._sizes.nbytes + self._pointers.nbytes)

        def __del__(self):
            try:
                self._bin_buffer_mmap._mmap.close()
            except Exception:
                pass

        @property
        def dtype(self):
            return self._dtype

        @property
        def sizes(self):
            return self._sizes

        @property
        def doc_idx(self):
            return self._doc_idx

        @lru_cache(maxsize=8)
        def __getitem__(self, i):
            return self._pointers[i], self._sizes[i]

        def __len__(self):
            return self._len

    def __init__(self, path, skip_warmup=False):
        super().__init__()
        self._path = self._index = self._bin_buffer = None
        self._do_init(path, skip_warmup)

    def __getstate__(self):
        return self._path

    def __setstate__(self, state):
        self._do_init(state, skip_warmup=True)

    def _do_init(self, path, skip_warmup):
        self._path = path
        self._index = self.Index(index_file_path(self._path), skip_warmup)
        if not skip_warmup:
            print_rank_0("    warming up data mmap file...")
            _warmup_mmap_file(data_file_path(self._path))
        self._bin_buffer_mmap = np.memmap(data_file_path(self._path), mode="r", order="C")
        self._bin_buffer = memoryview(self._bin_buffer_mmap)

    def __del__(self):
        try:
            self._bin_buffer_mmap._mmap.close()
        except Exception:
            pass

    def __len__(self):
        return len(self._index)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            ptr, size = self._index[int(idx)]
            return np.frombuffer(self._bin_buffer, dtype=self._index.dtype, count=size, offset=ptr)
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("Slices into indexed_dataset must be contiguous")
            ptr = self._index._pointers[start]
            sizes = self._index._sizes[idx]
            flat = np.frombuffer(self._bin_buffer, dtype=self._index.dtype, count=int(sum(sizes)), offset=ptr)
            return np.split(flat, list(accumulate(sizes))[:-1])
        raise TypeError(type(idx))

    def get(self, idx, offset=0, length=None):
        """Tokens ``[offset, offset+length)`` of sequence ``idx`` without touching the rest."""
        ptr, size = self._index[int(idx)]
        if length is None:
            length = size - offset
        ptr += offset * np.dtype(self._index.dtype).itemsize
        return np.frombuffer(self._bin_buffer, dtype=self._index.dtype, count=length, offset=ptr)

    @property
    def sizes(self):
        return self._index.sizes

    @property
    def doc_idx(self):
        return self._index.doc_idx

    def get_doc_idx(self):
        return self._index._doc_idx

    def set_doc_idx(self, doc_idx_):
        self._index._doc_idx = doc_idx_

    @property
    def supports_prefetch(self):
        return False

    @staticmethod
    def exists(path):
        return os.path.exists(index_file_path(path)) and os.path.exists(data_file_path(path))


class MMapIndexedDatasetBuilder:
    def __init__(self, out_file, dtype=np.int64):
        self._data_file = open(out_file, "wb")
        self._dtype = dtype
        self._sizes = []
        self._doc_idx = [0]

    def add_item(self, tensor):
        arr = np.array(tensor.numpy() if isinstance(tensor, torch.Tensor) else tensor, dtype=self._dtype)
        self._data_file.write(arr.tobytes(order="C"))
        self._sizes.append(arr.size)

    def add_doc(self, tensor, sizes):
        arr = np.array(tensor, dtype=self._dtype)
        self._data_file.write(arr.tobytes(order="C"))
        self._sizes.extend(sizes)
        self._doc_idx.append(len(self._sizes))

    def end_document(self):
        self._doc_idx.append(len(self._sizes))

    def merge_file_(self, another_file):
        index = MMapIndexedDataset.Index(index_file_path(another_file), skip_warmup=True)
        assert index.dtype == self._dtype
        offset = len(self._sizes)
        self._sizes.extend(index.sizes)
        self._doc_idx.extend((offset + index.doc_idx)[1:])
        with open(data_file_path(another_file), "rb") as f:
            shutil.copyfileobj(f, self._data_file)

    def finalize(self, index_file):
        self._data_file.close()
        with MMapIndexedDataset.Index.writer(index_file, self._dtype) as index:
            index.write(self._sizes, self._doc_idx)
