"""Kernel timeline of graph replays of the benchmarked step (torch.profiler / CUPTI): per kernel start offset,
duration, stream; gaps where no kernel of ours is running.  usage: python tools/timeline.py [bench.py flags] > out.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import exemplar_vae_b200 as E
    from exemplar_vae_b200.config import default_args
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    a = bench.parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    args = default_args(device="cuda", seed=0, **bench.model_kwargs(a))
    torch.manual_seed(0)
    model = E.importing_model(args)(args).to(dev)
    data = bench.synthetic_data(a)
    T = a.train_size
    dataset = torch.utils.data.TensorDataset(data, torch.arange(T).view(-1, 1), torch.zeros(T))
    opt = E.AdamNormGrad(model.parameters(), lr=5e-4)
    if world > 1:
        from exemplar_vae_b200 import distributed as D
        D.shard_bank(model, opt, dist.group.WORLD, shard=not a.approximate)
    cache = None
    if a.approximate:
        with torch.no_grad():
            cache = model.cache_z(dataset)
    graphable = (not a.approximate) or getattr(model, "knn_graph_capturable", False)
    step = E.GraphedTrainStep(model, opt, args, dataset, a.batch, use_graph=graphable and not a.no_graph, cache=cache)
    idx = torch.randint(0, T, (a.batch,))
    x, xi = data[idx].to(dev), idx.to(dev)
    for _ in range(5):
        step.step(x, xi)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step.step(x, xi)
        torch.cuda.synchronize()
    if rank != 0:
        torch.cuda.synchronize()
        os._exit(0)
    path = os.path.join(ROOT, "gpurun_out", "timeline_trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ks.sort(key=lambda e: e["ts"])
    # split into replays by the largest gaps
    gaps = sorted(((ks[i + 1]["ts"] - (ks[i]["ts"] + ks[i]["dur"]), i) for i in range(len(ks) - 1)), reverse=True)[:2]
    cuts = sorted(i for _, i in gaps)
    last = ks[cuts[-1] + 1:]
    t0 = last[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in last)
    print(f"# kernel timeline of one replay ({' '.join(sys.argv[1:]) or 'cfg2'}): span {end - t0:.1f} us, "
          f"{len(last)} device activities, sum of durations {sum(e['dur'] for e in last):.1f} us\n")
    # busy time = union of intervals
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in last)
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    idle = []
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            idle.append((cur_e - t0, s - cur_e))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print(f"device busy (union) {busy:.1f} us, idle inside the step {end - t0 - busy:.1f} us over {len(idle)} gaps\n")
    print("| start us | dur us | stream | kernel |\n|---|---|---|---|")
    for e in last:
        name = e["name"].replace("exvae::<unnamed>::", "").replace("void ", "")[:110]
        print(f"| {e['ts'] - t0:.1f} | {e['dur']:.1f} | {e.get('args', {}).get('stream', '?')} | `{name}` |")
    os.remove(path)
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
