"""Where does the host-buffer (e2e) path stop scaling?  Run under torch.distributed.run with N ranks (one per GPU): every rank
copies the bench workload's bytes (363 MB of PCM in, 106 MB of features out) between its own pinned host buffers and its GPU,
all ranks at the same time, and reports GB/s for H2D alone, D2H alone and both directions together; then the same with the
host buffers bound to the NUMA node the GPU hangs off (mbind before pinning), when the kernel lets us.  Rank 0 prints one JSON
line.  No kernels of ours run here: this measures the host fabric under the front-end's copy pattern."""
import ctypes
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

H2D, D2H = 363_070_656, 106_346_580


def gpu_numa_node(index):
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip().lower()
        bus = bus[4:] if bus.startswith("0000") and len(bus) > 12 else bus
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            return int(f.read())
    except Exception:
        return -1


def alloc_pinned(nbytes, node):
    """page-aligned host buffer, optionally bound to a NUMA node with mbind(MPOL_BIND) before it is touched, then pinned"""
    libc = ctypes.CDLL(None, use_errno=True)
    libc.mmap.restype = ctypes.c_void_p
    libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
    n = (nbytes + (2 << 20) - 1) & ~((2 << 20) - 1)
    p = libc.mmap(None, n, 3, 0x22, -1, 0)          # PROT_READ|WRITE, MAP_PRIVATE|MAP_ANONYMOUS
    bound = None
    if node >= 0:
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(237, ctypes.c_void_p(p), ctypes.c_ulong(n), 2, ctypes.byref(mask), ctypes.c_ulong(64), 0)   # mbind, MPOL_BIND
        bound = rc == 0
    ctypes.memset(p, 1, n)
    rt = torch.cuda.cudart()
    rc = rt.cudaHostRegister(p, n, 0)
    return p, n, bound, int(rc)


def bw(nbytes, fn, rep=5):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(rep):
        fn()
    torch.cuda.synchronize()
    return nbytes * rep / (time.perf_counter() - t0) / 1e9


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    node = gpu_numa_node(local)
    res = {"rank": rank, "gpu_numa_node": node, "cpus": sorted(os.sched_getaffinity(0))[:4] + ["..."], "n_cpus": len(os.sched_getaffinity(0))}
    d_in = torch.empty(H2D, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(D2H, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for label, nd in (("default", -1), ("bound_to_gpu_node", node)):
        if nd < 0 and label != "default":
            continue
        if label == "default":
            h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory(); h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
            cp_in = lambda: d_in.copy_(h_in, non_blocking=True)
            cp_out = lambda: h_out.copy_(d_out, non_blocking=True)
        else:
            pi, ni, bi, rci = alloc_pinned(H2D, nd); po, no, bo, rco = alloc_pinned(D2H, nd)
            res["mbind_ok"] = bool(bi and bo); res["host_register_rc"] = [rci, rco]
            if rci or rco:
                continue
            rt = torch.cuda.cudart()
            cp_in = lambda: rt.cudaMemcpyAsync(d_in.data_ptr(), pi, H2D, 1, torch.cuda.current_stream().cuda_stream)
            cp_out = lambda: rt.cudaMemcpyAsync(po, d_out.data_ptr(), D2H, 2, torch.cuda.current_stream().cuda_stream)

        def both():
            with torch.cuda.stream(s1):
                cp_in()
            with torch.cuda.stream(s2):
                cp_out()
        try:
            cp_in(); cp_out(); torch.cuda.synchronize()
            res[label] = {"h2d_gbs": bw(H2D, cp_in), "d2h_gbs": bw(D2H, cp_out), "both_gbs": bw(H2D + D2H, both)}
        except Exception as e:      # cudart binding without cudaMemcpyAsync etc.
            res[label] = {"error": repr(e)[:200]}
    allr = [None] * world
    dist.all_gather_object(allr, res)
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) if os.path.isdir("/sys/devices/system/node") else []
        print(json.dumps({"world": world, "ranks": allr, "numa_nodes": nodes, "topo": topo.splitlines()[:14]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
