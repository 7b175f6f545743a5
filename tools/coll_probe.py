"""Latency probe of the small collectives of the sharded step (8 ranks): NCCL under different algorithm hints and
torch symmetric-memory all-reduces.  torchrun ... tools/coll_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist


def timeit(fn, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tag = os.environ.get("PROBE_TAG", "default")
    res = {}
    for name, n in (("ar_1.5MB", 375000), ("ar_4.5MB", 1125000), ("ar_64KB", 16384)):
        t = torch.ones(n, device=dev)
        res[name] = timeit(lambda: dist.all_reduce(t, op=dist.ReduceOp.AVG))
    z = torch.ones(512 * 40, device=dev)
    zall = torch.empty(world * 512 * 40, device=dev)
    res["ag_80KB"] = timeit(lambda: dist.all_gather_into_tensor(zall, z))
    st = torch.ones(4096 * 4, device=dev)
    stall = torch.empty(world * 4096 * 4, device=dev)
    res["ag_64KBx8"] = timeit(lambda: dist.all_gather_into_tensor(stall, st))
    dz = torch.ones(4096 * 40, device=dev)
    dzo = torch.empty(512 * 40, device=dev)
    res["rs_640KB"] = timeit(lambda: dist.reduce_scatter_tensor(dzo, dz))
    if tag == "default":
        try:
            import torch.distributed._symmetric_memory as symm
            g = dist.group.WORLD
            for name, n in (("symm_ar_1.5MB", 375000), ("symm_ar_4.5MB", 1125000), ("symm_ar_64KB", 16384)):
                t = symm.empty(n, device=dev)
                t.fill_(1.0)
                symm.rendezvous(t, g.group_name)
                for opname in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_"):
                    try:
                        op = getattr(torch.ops.symm_mem, opname)
                        res[f"{name}:{opname}"] = timeit(lambda: op(t, "sum", g.group_name))
                    except Exception as ex:
                        res[f"{name}:{opname}"] = f"failed: {str(ex)[:80]}"
        except Exception as ex:
            res["symm"] = f"unavailable: {str(ex)[:120]}"
    if rank == 0:
        print(tag, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
