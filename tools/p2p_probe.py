"""Probe of the peer-memory plumbing the sequence-parallel path needs (run under torchrun, >= 2 ranks):
  1. torch.distributed._symmetric_memory: empty + rendezvous + peer pointers + barrier + a peer write;
  2. raw cudaMalloc + cudaIpc{Get,Open}MemHandle through libcudart (the fallback if 1 is unavailable).
Prints one line per check; exits 0 even when a check fails (the point is to learn which one works on the box)."""
import ctypes as C
import os
import sys
import traceback

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dev = torch.device("cuda", torch.cuda.current_device())
    peer = (rank + 1) % world
    print(f"[{rank}] can_access_peer({peer}) = {torch.cuda.can_device_access_peer(dev.index, peer)}", flush=True)

    # ---- 1. symmetric memory
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(1 << 20, dtype=torch.bfloat16, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        ptrs = list(hdl.buffer_ptrs)
        t.fill_(float(rank))
        hdl.barrier(channel=0)
        remote = hdl.get_buffer(peer, (1 << 20,), torch.bfloat16)
        remote[:16].fill_(100.0 + rank)          # peer store
        hdl.barrier(channel=0)
        torch.cuda.synchronize()
        src = (rank - 1) % world
        ok = bool((t[:16] == 100.0 + src).all()) and bool((t[16:] == float(rank)).all())
        print(f"[{rank}] symm_mem ok={ok} ptrs={[hex(p) for p in ptrs]} multicast_ptr={getattr(hdl, 'multicast_ptr', None)}", flush=True)
        # barrier latency
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            hdl.barrier(channel=0)
        s.record()
        for _ in range(100):
            hdl.barrier(channel=0)
        e.record()
        torch.cuda.synchronize()
        print(f"[{rank}] symm_mem barrier {s.elapsed_time(e) * 10:.1f} us", flush=True)
    except Exception:
        print(f"[{rank}] symm_mem FAILED:\n{traceback.format_exc()}", flush=True)

    # ---- 2. raw cudaIpc
    try:
        rt = C.CDLL("libcudart.so.12")
        ptr = C.c_void_p()
        nbytes = 2 << 20
        assert rt.cudaMalloc(C.byref(ptr), C.c_size_t(nbytes)) == 0
        handle = (C.c_byte * 64)()
        assert rt.cudaIpcGetMemHandle(handle, ptr) == 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        peers = []
        for q in range(world):
            if q == rank:
                peers.append(ptr.value)
                continue
            h = (C.c_byte * 64).from_buffer_copy(handles[q])
            out = C.c_void_p()
            class _H(C.Structure):
                _fields_ = [("reserved", C.c_byte * 64)]
            rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), _H, C.c_uint]
            hs = _H()
            C.memmove(C.byref(hs), h, 64)
            rc = rt.cudaIpcOpenMemHandle(C.byref(out), hs, 1)
            assert rc == 0, f"cudaIpcOpenMemHandle rc={rc}"
            peers.append(out.value)

        class _Raw:
            def __init__(self, p, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i2", "data": (p, False), "version": 3}

        loc = torch.as_tensor(_Raw(peers[rank], nbytes // 2), device=dev)
        rem = torch.as_tensor(_Raw(peers[peer], nbytes // 2), device=dev)
        loc.fill_(rank)
        torch.cuda.synchronize()
        dist.barrier()
        rem[:16].fill_(100 + rank)
        torch.cuda.synchronize()
        dist.barrier()
        src = (rank - 1) % world
        ok = bool((loc[:16] == 100 + src).all()) and bool((loc[16:] == rank).all())
        print(f"[{rank}] cudaIpc ok={ok}", flush=True)
    except Exception:
        print(f"[{rank}] cudaIpc FAILED:\n{traceback.format_exc()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
