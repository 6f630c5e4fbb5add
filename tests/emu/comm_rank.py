"""One emulated rank of a peer-memory collective (csrc/comm.cu built for the host by host_build.py).

    python comm_rank.py <lib.so> <dir> <rank> <world> <n> <mode> <epoch> <num_ctas> <ranks that take part, e.g. 0,1>

The "GPUs" are processes; symmetric memory is a set of files in <dir> that every rank maps (``buf<r>.bin``: the fp32
bucket of rank r, ``pad<r>.bin``: its int32 signal pad), so the cross-rank flag protocol of the real kernels -- publish
/ spin with acquire-release, the last-CTA counter, the bounded spin with its error flag -- runs between really
concurrent peers.  mode: peer_ar | peer_rs (16-byte loads from every peer) | nvls_ar | nvls_rs (multimem emulated over
the mapped copies) | barrier.  No torch import: start-up stays cheap."""
import ctypes
import os
import sys
import time

import numpy as np


def main():
    so, d, rank, world, n, mode, epoch, ctas, present = sys.argv[1:10]
    rank, world, n, epoch, ctas = int(rank), int(world), int(n), int(epoch), int(ctas)
    lib = ctypes.CDLL(so)
    bufs = [np.memmap(os.path.join(d, f"buf{r}.bin"), dtype=np.float32, mode="r+", shape=(n,)) for r in range(world)]
    pads = [np.memmap(os.path.join(d, f"pad{r}.bin"), dtype=np.int32, mode="r+", shape=(64,)) for r in range(world)]
    addr = lambda a: a.ctypes.data
    peer = (ctypes.c_longlong * world)(*[addr(b) for b in bufs])
    pad_peer = (ctypes.c_longlong * world)(*[addr(p) for p in pads])
    local, pad_local = ctypes.c_void_p(addr(bufs[rank])), ctypes.c_void_p(addr(pads[rank]))
    scale = ctypes.c_float(1.0 / world)
    # rendezvous of the ranks that take part (interpreter start-up must not count against the kernels' bounded spins)
    open(os.path.join(d, f"ready{epoch}_{rank}"), "w").close()
    while not all(os.path.exists(os.path.join(d, f"ready{epoch}_{r}")) for r in map(int, present.split(","))):
        time.sleep(0.001)
    if mode in ("peer_ar", "peer_rs"):
        rc = lib.mlb_dp_reduce(int(mode == "peer_rs"), local, peer, pad_local, pad_peer, ctypes.c_longlong(n), rank, world,
                               epoch, scale, ctas, None)
    elif mode in ("nvls_ar", "nvls_rs"):
        mc = np.zeros(n, dtype=np.float32)                       # only its ADDRESS range is used (the multicast VA)
        lib.emu_register_multicast(ctypes.c_void_p(addr(mc)), peer, world, ctypes.c_longlong(n))
        rc = lib.mlb_dp_reduce_nvls(int(mode == "nvls_rs"), local, ctypes.c_void_p(addr(mc)), pad_local, pad_peer,
                                    ctypes.c_longlong(n), rank, world, epoch, scale, ctas, None)
    else:
        rc = lib.mlb_peer_barrier(pad_local, pad_peer, rank, world, epoch, 48, None)
    for a in bufs + pads:
        a.flush()
    return rc


if __name__ == "__main__":
    sys.exit(main())
