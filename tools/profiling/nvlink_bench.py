"""NVLink transfer micro-benchmark: GB/s per GPU for the instruction paths a fused GEMM+collective kernel can use.

  torchrun --nproc-per-node N tools/profiling/nvlink_bench.py [MB_per_peer]     (rank 0 prints one JSON line per case)

Every rank runs the same transfer at the same time (device-timed, max over ranks):
  pattern "ring": all bytes go to / come from rank+1;  "all": an equal share to / from each of the N-1 peers
  (concurrent streams), which is what all-gather / reduce-scatter traffic looks like.
Modes: see csrc/p2p_bench.cu (ld/st, bulk copy pipeline, bulk row stores = 2-D TMA store box pattern, multimem.st).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    import torch.distributed._symmetric_memory as symm
    from megatron_llm_b200.ops import _ext
    mod = _ext.load()
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    per_peer = mb << 20
    total = per_peer * max(world - 1, 1)
    dev = torch.device("cuda", rank)
    src = symm.empty(total, dtype=torch.uint8, device=dev)
    dst = symm.empty(total, dtype=torch.uint8, device=dev)
    h_src = symm.rendezvous(src, dist.group.WORLD)
    h_dst = symm.rendezvous(dst, dist.group.WORLD)
    src.random_(0, 255)
    src_ptrs = [int(p) for p in h_src.buffer_ptrs]
    dst_ptrs = [int(p) for p in h_dst.buffer_ptrs]
    mc_dst = int(getattr(h_dst, "multicast_ptr", 0) or 0)
    streams = [torch.cuda.Stream() for _ in range(max(world - 1, 1))]
    torch.cuda.synchronize()
    dist.barrier()

    def run(direction, pattern, mode, ctas, piece=32768, stages=6, row=0, stride=0):
        """direction 'pull': src = peer, dst = local; 'push': src = local, dst = peer."""
        peers = [(rank + 1) % world] if pattern == "ring" else [(rank + i) % world for i in range(1, world)]
        nbytes = total if pattern == "ring" else per_peer
        if mode == 2:          # rows are `stride` apart at the destination: stay inside the buffer
            nbytes = (nbytes * row // stride) // piece * piece
        share = max(ctas // len(peers), 1)

        def once():
            main_s = torch.cuda.current_stream()
            for i, p in enumerate(peers):
                st = streams[i]
                st.wait_stream(main_s)
                off = 0 if pattern == "ring" else i * per_peer
                s_ptr = (src_ptrs[p] if direction == "pull" else src_ptrs[rank]) + off
                d_ptr = (dst_ptrs[rank] if direction == "pull" else dst_ptrs[p]) + off
                with torch.cuda.stream(st):
                    mod.p2p_bench(mode, s_ptr, d_ptr, nbytes, share, piece, stages, row, stride)
            for st in streams[:len(peers)]:
                main_s.wait_stream(st)

        for _ in range(2):
            once()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            once()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / 5], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        moved = nbytes * len(peers)
        res = {"world": world, "direction": direction, "pattern": pattern,
               "mode": {0: "ldst", 1: "bulk", 2: "bulk_rows", 3: "multimem"}[mode], "ctas_total": share * len(peers),
               "piece": piece, "stages": stages, "row_bytes": row, "MB": moved >> 20, "us": t.item() * 1e3,
               "GBps_per_gpu": moved / t.item() / 1e6}
        if rank == 0:
            print(json.dumps(res), flush=True)

    patterns = ["ring"] + (["all"] if world > 2 else [])
    for pattern in patterns:
        for direction in ("pull", "push"):
            for ctas in (8, 16, 32, 64, 128):
                run(direction, pattern, 0, ctas)
            for ctas in (4, 8, 16, 32):
                run(direction, pattern, 1, ctas, piece=32768, stages=6)
            run(direction, pattern, 1, 16, piece=16384, stages=12)
            run(direction, pattern, 1, 16, piece=65536, stages=3)
        for row in (128, 256, 512, 1024, 4096):
            run("push", pattern, 2, 16, piece=32768, stages=6, row=row, stride=2 * row)
            run("push", pattern, 2, 32, piece=32768, stages=6, row=row, stride=2 * row)
    if mc_dst:
        # one multicast store reaches every rank's dst: per-GPU egress = bytes, ingress = world x bytes
        for ctas in (8, 16, 32, 64):
            for _ in range(2):
                mod.p2p_bench(3, src_ptrs[rank], mc_dst, per_peer, ctas, 0, 0, 0, 0)
            torch.cuda.synchronize()
            dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                mod.p2p_bench(3, src_ptrs[rank], mc_dst, per_peer, ctas, 0, 0, 0, 0)
            e.record()
            torch.cuda.synchronize()
            t = torch.tensor([s.elapsed_time(e) / 5], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({"world": world, "mode": "multimem", "ctas_total": ctas, "MB_sent": per_peer >> 20,
                                  "us": t.item() * 1e3, "egress_GBps": per_peer / t.item() / 1e6,
                                  "ingress_GBps": per_peer * world / t.item() / 1e6}), flush=True)
    elif rank == 0:
        print(json.dumps({"multimem": "no multicast pointer on this symmetric allocation"}), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


main()
