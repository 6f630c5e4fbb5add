"""Symmetric-memory communicators: peer-mapped buffers + signal pads for the fused collective kernels.

The reference only ever calls NCCL through ``torch.distributed``.  Here every tensor-parallel /
data-parallel group that lives inside one NVSwitch domain gets a small symmetric heap (allocated and exchanged with
``torch.distributed._symmetric_memory``: VMM allocations mapped into all peers) and the hand-written kernels do the
transfers themselves:

* :class:`TPCommunicator`  -- ``ag_gemm`` (all-gather -> GEMM) and ``gemm_rs`` (GEMM -> reduce-scatter), see
  ``csrc/gemm_sm100.cuh`` MODE_AG_GEMM / MODE_GEMM_RS.
* :class:`DPCommunicator`  -- bucketed two-shot all-reduce / reduce-scatter of the fp32 gradient buffer fused with the
  1/DP scale (``csrc/comm.cu``), launched from the backward hooks on a side stream.

Everything degrades to the NCCL path (the checked fallback) if symmetric memory cannot be set up.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from ..ops import _ext
from . import fused_tp
from . import state as ps

_PAD_INTS = 64


def _symm():
    import torch.distributed._symmetric_memory as symm_mem
    return symm_mem


def _alloc_symmetric(numel: int, dtype, device, group):
    """(local tensor, handle) for a buffer mapped into every rank of ``group``."""
    sm = _symm()
    t = sm.empty(numel, dtype=dtype, device=device)
    hdl = sm.rendezvous(t, group)
    return t, hdl


class LoopbackWorld:
    """``world`` virtual ranks on ONE device (tests / single-GPU leases): the "symmetric" buffers are ordinary
    allocations of the same process, so every rank's kernels see the others' buffers through plain device pointers,
    and the rank kernels are launched on separate streams with ``sms`` SMs each so that all of them are resident at the
    same time (they spin on each other's flags).  Exercises the whole flag / epoch / slot protocol and the address
    arithmetic of the fused kernels without NVLink."""

    def __init__(self, world: int, sms: Optional[int] = None, device=None):
        self.world = world
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        total = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.sms = sms if sms is not None else max(2, ((total - 4) // world) & ~1)
        self._bufs = {}

    def alloc(self, name: str, numel: int, dtype):
        """-> list of ``world`` tensors (index = rank); created on first use, shared by all virtual ranks."""
        if name not in self._bufs:
            self._bufs[name] = [torch.zeros(numel, dtype=dtype, device=self.device) for _ in range(self.world)]
        return self._bufs[name]


_COMMUNICATORS = []      # every live TP / DP communicator of this process (polled by check_timeouts)
_POLL = {"buf": None, "event": None}


def check_timeouts() -> None:
    """Called once per training step: a bounded spin-wait inside a fused kernel that gave up (lost / hung peer) leaves
    ``PAD_ERROR`` in the communicator's signal pad and the kernel's results are garbage.  The flags are fetched with an
    asynchronous copy and examined when it has landed (no host sync on the step path), so the failure surfaces at most
    one step late -- as an exception, never as silently wrong training."""
    if not _COMMUNICATORS or torch.cuda.is_current_stream_capturing():
        return
    if os.environ.get("MLB200_TIMEOUT_FATAL", "1") == "0":      # (bench.py checks the flags itself and falls back)
        return
    ev = _POLL["event"]
    if ev is not None and ev.query():
        bad = [i for i, v in enumerate(_POLL["buf"][:len(_COMMUNICATORS)].tolist()) if v != 0]
        _POLL["event"] = None
        if bad:
            names = ", ".join(type(_COMMUNICATORS[i]).__name__ for i in bad)
            raise RuntimeError(f"peer-memory collective timed out waiting for a peer ({names}): a rank is lost or hung; "
                               "results of the affected kernels are invalid")
    if _POLL["event"] is None:
        if _POLL["buf"] is None or _POLL["buf"].numel() < len(_COMMUNICATORS):
            _POLL["buf"] = torch.zeros(max(8, len(_COMMUNICATORS)), dtype=torch.int32).pin_memory()
        for i, c in enumerate(_COMMUNICATORS):
            _POLL["buf"][i:i + 1].copy_(c.pad[32:33], non_blocking=True)
        e = torch.cuda.Event()
        e.record()
        _POLL["event"] = e


class TPCommunicator:
    """Fused GEMM+collective for one tensor-parallel group.

    Buffers (per rank, symmetric):
      xs   [max_rows_per_rank * max_k]            bf16  -- this rank's published activation shard (all-gather source)
      rs   [2][world][max_rows_per_rank * max_n]  bf16  -- receive slots for reduce-scatter tiles (2 epoch parities)
      pad  [64] int32                                    -- ready / ack / arrived / free signals
    """

    def __init__(self, group, max_rows_per_rank: int, max_k: int, max_n: int, num_comm_ctas: int = 8,
                 loopback: Optional[LoopbackWorld] = None, loopback_rank: int = 0, all_reduce_n: int = 0,
                 nvls_ag_k: int = 0):
        self.group = group
        self.loopback = loopback
        if loopback is not None:
            self.world, self.rank, self.device = loopback.world, loopback_rank, loopback.device
            self.sms = loopback.sms          # SMs this virtual rank's kernels may occupy
        else:
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
            self.device = torch.device("cuda", torch.cuda.current_device())
            self.sms = 0                     # 0 = all SMs of the device
        self.max_rows, self.max_k, self.max_n = max_rows_per_rank, max_k, max_n
        self.num_comm_ctas = num_comm_ctas
        self.enabled = False
        self.mod = _ext.load()
        self.xs, self.xs_ptrs = self._symmetric("xs", max_rows_per_rank * max_k, torch.bfloat16)
        self.rs, self.rs_ptrs = self._symmetric("rs", 2 * self.world * max_rows_per_rank * max_n, torch.bfloat16)
        self.pad, self.pad_ptrs = self._symmetric("pad", _PAD_INTS, torch.int32)
        self.pad.zero_()
        # GEMM -> all-reduce (non-sequence-parallel layers): two parities of a [world * rows, all_reduce_n] output buffer
        self.ar_n = all_reduce_n
        if all_reduce_n > 0:
            self.ar, self.ar_ptrs = self._symmetric("ar", 2 * self.world * max_rows_per_rank * all_reduce_n,
                                                    torch.bfloat16)
        # NVLS all-gather (MLB200_AG_NVLS=1): two parities of a multicast-mapped gather buffer [world * rows, nvls_ag_k]
        # + per-chunk arrival flags; used by the 2-CTA fused kernel for operands with K <= nvls_ag_k
        self.nvls_ag = False
        self.ag_k = nvls_ag_k
        if nvls_ag_k > 0 and loopback is None and os.environ.get("MLB200_AG_NVLS", "0") == "1":
            chunks = max(self.world * max_rows_per_rank // 128, 1)
            self.agbuf, self.agbuf_ptrs = self._symmetric("agbuf", 2 * self.world * max_rows_per_rank * nvls_ag_k,
                                                          torch.bfloat16)
            self.agflag, self.agflag_ptrs = self._symmetric("agflag", 2 * chunks, torch.int32)
            self.agflag.zero_()
            self.agbuf_mc = self._mc.get("agbuf", 0)
            self.ag_done = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.push_ctas = int(os.environ.get("MLB200_AG_PUSH_CTAS", "8"))
            self.nvls_ag = self.agbuf_mc != 0
        max_chunks = self.world * max_rows_per_rank // 128
        self.chunk_flags = torch.zeros(max(max_chunks, 1), dtype=torch.int32, device=self.device)
        self.read_counters = torch.zeros(8, dtype=torch.int32, device=self.device)   # pullers done per peer (self-resetting)
        self.reduce_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.ag_epoch = 0
        self.rs_epoch = 0
        self.rs_arrived_total = 0
        # offsets that calls captured in a CUDA graph add to their (frozen) epoch arguments -- see replay_offsets()
        self.state = torch.zeros(4, dtype=torch.int32, device=self.device)
        # exposed-communication accounting written by the kernels (ns, averaged over the CTAs of each launch)
        self.stats = torch.zeros(2, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize()
        if loopback is None:
            dist.barrier(group=group)
            _COMMUNICATORS.append(self)
        self.enabled = True

    def _symmetric(self, name: str, numel: int, dtype):
        """(this rank's buffer, every rank's device pointer to its copy)"""
        if self.loopback is not None:
            bufs = self.loopback.alloc(name, numel, dtype)
            return bufs[self.rank], [int(t.data_ptr()) for t in bufs]
        t, hdl = _alloc_symmetric(numel, dtype, self.device, self.group)
        self._handles = getattr(self, "_handles", []) + [hdl]       # keep the peer mappings alive
        # NVSwitch multicast address of the same allocation (0 when the platform has no NVLS)
        self._mc = getattr(self, "_mc", {})
        self._mc[name] = int(getattr(hdl, "multicast_ptr", 0) or 0)
        return t, [int(p) for p in hdl.buffer_ptrs]

    @classmethod
    def loopback_group(cls, world: int, max_rows_per_rank: int, max_k: int, max_n: int, num_comm_ctas: int = 4,
                       sms: Optional[int] = None, all_reduce_n: int = 0):
        """``world`` communicators that talk to each other on the current device (see :class:`LoopbackWorld`)."""
        lw = LoopbackWorld(world, sms)
        return [cls(None, max_rows_per_rank, max_k, max_n, num_comm_ctas, loopback=lw, loopback_rank=r,
                    all_reduce_n=all_reduce_n) for r in range(world)]

    # -------------------------------------------------------------------------------------------
    def supports(self, rows_per_rank: int, k: int, n: int) -> bool:
        return (self.enabled and rows_per_rank % 128 == 0 and rows_per_rank <= self.max_rows and k % 8 == 0
                and n % 8 == 0)

    def ag_gemm(self, x_shard: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False, out=None,
                keep: bool = True):
        """x_shard [m, ..., K] (this rank's rows) -> (out [world*m*..., N], gathered [world*m, ..., K]).

        ``keep`` is accepted for interface stability (a transport that gathers into communicator-owned memory would
        have to copy when the caller keeps ``gathered`` for backward); this one gathers into a fresh tensor."""
        lead = x_shard.shape[:-1]
        K = x_shard.size(-1)
        x2d = x_shard.reshape(-1, K)
        m = x2d.size(0)
        N = weight.size(1) if transposed_weight else weight.size(0)
        assert m <= self.max_rows and K <= self.max_k and m % 128 == 0, (m, K, self.max_rows, self.max_k)
        if self.nvls_ag and K <= self.ag_k and m % 256 == 0 and N >= 256:
            res = self._ag_gemm_nvls(x2d, weight, transposed_weight, out, keep, lead, m, K, N)
            if res is not None:
                return res
        # publish my shard (stream-ordered before the kernel; the previous call's kernel only retired after every
        # peer had acknowledged reading the old content)
        gathered = torch.empty((self.world * m, K), dtype=torch.bfloat16, device=self.device)
        # one pass over my rows: into the symmetric buffer the peers pull from, and into their place in `gathered`
        # (the kernel's puller CTAs only move the remote shards)
        mine = gathered[self.rank * m: (self.rank + 1) * m]
        if x2d.is_contiguous() and x2d.data_ptr() % 16 == 0:
            self.mod.comm_copy2(x2d, self.xs[: m * K].view(m, K), mine)
            _ext.count()
        else:
            self.xs[: m * K].view(m, K).copy_(x2d)
            mine.copy_(x2d)
        if out is None:
            out = torch.empty((self.world * m, N), dtype=torch.bfloat16, device=self.device)
        self.ag_epoch += 1
        w = weight if weight.stride(-1) == 1 else weight.contiguous()
        self.mod.fused_ag_gemm(gathered, w, out, transposed_weight, self.xs_ptrs, m, self.chunk_flags,
                               self.read_counters, self.pad_ptrs[self.rank], self.pad_ptrs, self.rank, self.world,
                               self.ag_epoch, self.num_comm_ctas, self.sms, self._state_ptr(), self.stats.data_ptr())
        _ext.count()
        return out, gathered.view(self.world * lead[0], *lead[1:], K)

    def _ag_gemm_nvls(self, x2d, weight, transposed_weight, out, keep, lead, m, K, N):
        """NVLS transport: pusher CTAs multicast this rank's shard into every rank's gather buffer (see
        ``ag_pusher_nvls`` in csrc/gemm_sm100.cuh); ``gathered`` is a view of that buffer unless the caller keeps it."""
        epoch = self.ag_epoch + 1
        parity = epoch % 2
        buf_elems = self.world * self.max_rows * self.ag_k           # one parity of the gather buffer
        flag_ints = self.agflag.numel() // 2
        view = self.agbuf[parity * buf_elems: parity * buf_elems + self.world * m * K].view(self.world * m, K)
        mc_dst = self.agbuf_mc + 2 * parity * buf_elems
        flags = [p + 4 * parity * flag_ints for p in self.agflag_ptrs]
        if out is None:
            out = torch.empty((self.world * m, N), dtype=torch.bfloat16, device=self.device)
        x = x2d.contiguous()
        if x.data_ptr() % 16:
            x = x.clone()
        w = weight if weight.stride(-1) == 1 else weight.contiguous()
        ok = self.mod.fused_ag_gemm_nvls(view, x, w, out, transposed_weight, mc_dst, flags, self.ag_done, m,
                                         self.pad_ptrs[self.rank], self.pad_ptrs, self.rank, self.world, epoch,
                                         self.push_ctas, self.sms, self._state_ptr(), self.stats.data_ptr())
        if not ok:
            return None
        self.ag_epoch = epoch
        _ext.count()
        gathered = view.clone() if keep else view
        return out, gathered.view(self.world * lead[0], *lead[1:], K)

    def gemm_rs(self, x2d: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False):
        """reduce_scatter_rows(x2d [M, K] @ W^T or W) -> [M/world, N]."""
        return self._gemm_rs(x2d, weight, transposed_weight, all_reduce=False)

    def gemm_ar(self, x2d: torch.Tensor, weight: torch.Tensor, transposed_weight: bool = False, keep: bool = True):
        """all_reduce(x2d [M, K] @ W^T or W) -> [M, N]: the reduce-scatter kernel whose reduction phase stores every
        rank's reduced row slice into ALL ranks' symmetric output buffers (GEMM -> reduce-scatter -> all-gather in one
        launch).  ``keep=False``: the caller consumes the result before the next-but-one call on this communicator and
        gets a view of the communicator's buffer instead of a copy."""
        return self._gemm_rs(x2d, weight, transposed_weight, all_reduce=True, keep=keep)

    def _gemm_rs(self, x2d, weight, transposed_weight, all_reduce, keep=True):
        M, K = x2d.shape
        N = weight.size(1) if transposed_weight else weight.size(0)
        m = M // self.world
        assert m <= self.max_rows and N <= self.max_n and m % 128 == 0 and m * self.world == M, (M, N)
        self.rs_epoch += 1
        parity = self.rs_epoch % 2
        slot_elems = m * N
        # my receive slot on rank d for this parity: rs_d[parity][src=self.rank]
        base = (parity * self.world) * self.max_rows * self.max_n
        rs_dst = [self.rs_ptrs[d] + 2 * (base + self.rank * slot_elems) for d in range(self.world)]
        rs_slots = self.rs_ptrs[self.rank] + 2 * base
        ar_dst = []
        if all_reduce:
            assert 0 < N <= self.ar_n, "communicator was built without (large enough) all-reduce buffers"
            ar_base = parity * self.world * self.max_rows * self.ar_n
            ar_dst = [p + 2 * ar_base for p in self.ar_ptrs]
            full = self.ar[ar_base: ar_base + M * N].view(M, N)
            out = full[self.rank * m: (self.rank + 1) * m]          # (unused by the kernel in this mode)
        else:
            out = torch.empty((m, N), dtype=torch.bfloat16, device=self.device)
        tiles_per_dst = (m // 128) * self._num_n_tiles(M, N)      # arrivals if the 1-CTA kernel is chosen
        x = x2d if (x2d.stride(1) == 1 and x2d.stride(0) % 8 == 0) else x2d.contiguous()
        w = weight if weight.stride(-1) == 1 else weight.contiguous()
        self.rs_arrived_total = self.mod.fused_gemm_rs(
            x, w, out, transposed_weight, rs_dst, rs_slots, m, self.rs_arrived_total, tiles_per_dst,
            self.reduce_counter, self.pad_ptrs[self.rank], self.pad_ptrs, self.rank, self.world, self.rs_epoch,
            self.sms, self._state_ptr(), ar_dst, self.stats.data_ptr())
        _ext.count()
        if all_reduce:
            return full.clone() if keep else full
        return out

    # ------------------------------------------------------------------- CUDA-graph support
    # Epochs / arrival targets are host-side counters passed as kernel arguments.  Calls issued while a stream is
    # capturing keep the values of the capture pass forever, so they additionally read an offset from ``self.state``
    # (device memory); eager calls pass no state pointer and use their arguments as they are.
    def _state_ptr(self) -> int:
        return self.state.data_ptr() if torch.cuda.is_current_stream_capturing() else 0

    def counters(self):
        return (self.ag_epoch, self.rs_epoch, self.rs_arrived_total)

    def end_capture(self, before):
        """Called right after a capture pass that started at ``counters() == before``: the captured calls did not
        run, so rewind, and return what one replay advances the counters by."""
        adv = tuple(now - b for now, b in zip(self.counters(), before))
        self.ag_epoch, self.rs_epoch, self.rs_arrived_total = before
        return adv

    def begin_replay(self, before, advance):
        """Enqueue the offset update that makes a replay of a graph captured at ``before`` continue the live
        sequence, and advance the host counters past it.  (The receive-slot parity of a captured reduce-scatter is
        frozen too, so the rs epoch offset is kept even by skipping one epoch number when needed -- the slot-free
        handshake only needs monotonic epochs.)"""
        if (self.rs_epoch - before[1]) % 2:
            self.rs_epoch += 1
        if getattr(self, "nvls_ag", False) and (self.ag_epoch - before[0]) % 2:     # (NVLS gather buffer parity: frozen too)
            self.ag_epoch += 1
        self.mod.comm_set_state(self.state, self.ag_epoch - before[0], self.rs_epoch - before[1],
                                self.rs_arrived_total - before[2])
        self.ag_epoch += advance[0]
        self.rs_epoch += advance[1]
        self.rs_arrived_total += advance[2]

    def _num_n_tiles(self, M: int, N: int) -> int:
        """Mirror of ``pick_block_n`` in csrc/gemm_sm100.cu (the arrival counters count output tiles)."""
        sms = self.sms or self.mod.num_sms()

        def cost(bn):
            tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
            return ((tiles + sms - 1) // sms) * bn / (1.0 if bn == 256 else 0.78)

        bn = 256 if cost(256) <= cost(128) else 128
        return (N + bn - 1) // bn

    def error_flag(self) -> int:
        return int(self.pad[32].item())

    def exposed_ms(self):
        """(all-gather wait, reduce-scatter tail) in ms accumulated since the communicator was built: time the fused
        kernels could not hide behind their GEMM tiles (device ``%globaltimer`` stamps; host sync)."""
        ag, rs = self.stats.tolist()
        return ag / 1e6, rs / 1e6


class DPCommunicator:
    """Peer-memory gradient reduction for one data-parallel group: the whole fp32 grad buffer is symmetric."""

    def __init__(self, group, numel_padded: int, num_ctas: int = 16, loopback: Optional[LoopbackWorld] = None,
                 loopback_rank: int = 0):
        self.group = group
        self.loopback = loopback
        if loopback is not None:
            self.world, self.rank, self.device = loopback.world, loopback_rank, loopback.device
        else:
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.mod = _ext.load()
        self.num_ctas = num_ctas
        self.enabled = False
        self.buffer, self.buf_ptrs = TPCommunicator._symmetric(self, "dp_buffer", numel_padded, torch.float32)
        self.buffer.zero_()
        self.pad, self.pad_ptrs = TPCommunicator._symmetric(self, "dp_pad", _PAD_INTS, torch.int32)
        self.pad.zero_()
        self.epoch = 0
        self.stream = torch.cuda.Stream(priority=-1)
        # NVLS: reduce inside the switch (multimem.ld_reduce) and write back by multicast when the grad buffer has a
        # multicast mapping; MLB200_DP_NVLS=0 keeps the peer-pointer kernel
        self.buf_mc = getattr(self, "_mc", {}).get("dp_buffer", 0) if loopback is None else 0
        self.use_nvls = self.buf_mc != 0 and os.environ.get("MLB200_DP_NVLS", "1") == "1"
        # measured on 4 x B200, 1 GB fp32 bucket (profiles/dp_reduce_n4_r2.jsonl): the NVLS kernel is link-bound with
        # 16 CTAs (all-reduce 2.35 ms vs NCCL 2.57 ms); the peer-pointer kernel scales with its CTA count (16: 14.9 ms,
        # 64: 4.0 ms) -- give it 64 when it has to be used
        if loopback is None and not self.use_nvls:
            self.num_ctas = max(self.num_ctas, 64)
        torch.cuda.synchronize()
        if loopback is None:
            dist.barrier(group=group)
            _COMMUNICATORS.append(self)
        self.enabled = True

    def error_flag(self) -> int:
        return int(self.pad[32].item())

    # ---- ZeRO-1 parameter all-gather fused into the optimizer kernel -----------------------------------------------
    def attach_param_buffer(self, numel: int, dtype):
        """Symmetric twin of the DDP wrapper's flat 16-bit parameter buffer: every rank's AdamW kernel stores its
        updated shard into all of them (``adamw_flat(..., p16_peers)``), :meth:`params_barrier` closes the step."""
        self.pbuf, self.pbuf_ptrs = TPCommunicator._symmetric(self, "dp_params", numel, dtype)
        self.param_epoch = 0
        return self.pbuf

    def param_peer_ptrs(self, start_elem: int):
        """Device pointers of element ``start_elem`` of every rank's parameter buffer (own rank first)."""
        es = self.pbuf.element_size()
        order = [(self.rank + k) % self.world for k in range(self.world)]
        return [self.pbuf_ptrs[d] + es * start_elem for d in order]

    def params_barrier(self):
        """All ranks' optimizer kernels have finished storing into everybody's parameter buffer."""
        self.param_epoch += 1
        self.mod.peer_barrier(self.pad_ptrs[self.rank], self.pad_ptrs, self.rank, self.world, self.param_epoch, 16)
        _ext.count()

    @classmethod
    def loopback_group(cls, world: int, numel_padded: int, num_ctas: int = 16):
        lw = LoopbackWorld(world)
        return [cls(None, numel_padded, num_ctas, loopback=lw, loopback_rank=r) for r in range(world)]

    def reduce_bucket(self, view: torch.Tensor, start: int, numel_padded: int, reduce_scatter: bool):
        """``view`` = buffer[start:end] (a slice of the symmetric grad buffer).  Returns an object with ``wait()``."""
        n = view.numel()
        assert n % (self.world * 4) == 0
        self.epoch += 1
        ev = torch.cuda.Event()
        ev.record()                       # the bucket's last gradient was produced on the current stream
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            if self.use_nvls:
                self.mod.dp_reduce_nvls(view, self.buf_mc + 4 * start, self.pad_ptrs[self.rank], self.pad_ptrs,
                                        self.rank, self.world, self.epoch, 1.0 / self.world, reduce_scatter,
                                        self.num_ctas)
            else:
                peers = [p + 4 * start for p in self.buf_ptrs]
                self.mod.dp_reduce(view, peers, self.pad_ptrs[self.rank], self.pad_ptrs, self.rank, self.world,
                                   self.epoch, 1.0 / self.world, reduce_scatter, self.num_ctas)
            done = torch.cuda.Event()
            done.record()
        _ext.count()

        class _Handle:
            def wait(self_inner):
                torch.cuda.current_stream().wait_event(done)
        return _Handle()


def bind_tp_communicator(args) -> Optional[TPCommunicator]:
    """Create the TP group's communicator sized for this model and route ColumnParallel/RowParallel through it."""
    if os.environ.get("MLB200_FUSED_TP", "1") == "0":
        return None
    tp = ps.get_tensor_model_parallel_world_size()
    rows = args.seq_length * args.micro_batch_size // tp
    if rows % 128 != 0:
        return None
    ffn = args.ffn_hidden_size * (2 if args.glu_activation else 1)
    max_k = max(args.hidden_size, args.ffn_hidden_size // tp, ffn // tp)
    vocab_shard = getattr(args, "padded_vocab_size", 0) // tp
    max_n = max(args.hidden_size, ffn // tp, vocab_shard)
    comm = TPCommunicator(ps.get_tensor_model_parallel_group(), rows, max_k, max_n,
                          num_comm_ctas=int(os.environ.get("MLB200_AG_CTAS", "32")),   # upper bound: picked per shape
                          # without sequence parallelism the Row forward / Column dgrad end in an all-reduce of [s*b, h]
                          all_reduce_n=0 if args.sequence_parallel else args.hidden_size,
                          nvls_ag_k=args.hidden_size)       # every all-gathered operand has K = hidden size
    fused_tp.bind(comm)
    return comm


def bind_dp_communicator(ddp_module, group=None) -> Optional[DPCommunicator]:
    """Move a LocalDDP grad buffer into symmetric memory and reduce its buckets with the peer-memory kernel."""
    if os.environ.get("MLB200_FUSED_DP", "1") == "0":
        return None
    group = group or ps.get_data_parallel_group()
    if dist.get_world_size(group) == 1:
        return None
    bufs = ddp_module.grad_buffers()
    if torch.float32 not in bufs:
        return None
    mb = bufs[torch.float32]
    comm = DPCommunicator(group, mb.numel_padded)
    ddp_module.rehome_grad_buffer(torch.float32, comm.buffer)
    ddp_module.bind_symmetric_communicator(comm)
    # ZeRO-1: the 16-bit weights live in symmetric memory too, so the optimizer kernel can all-gather them as it casts
    if getattr(ddp_module, "use_distributed_optimizer", False) and os.environ.get("MLB200_FUSED_PARAM_GATHER", "1") == "1":
        pbufs = ddp_module.param_buffers().get(torch.float32, {})
        for pdtype, pb in list(pbufs.items()):
            if pdtype in (torch.bfloat16, torch.float16):
                ddp_module.rehome_param_buffer(torch.float32, pdtype, comm.attach_param_buffer(pb.numel(), pdtype))
    return comm
