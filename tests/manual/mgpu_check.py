"""Multi-GPU parity check (run under torchrun on N GPUs): a training step with the exemplar bank
range-sharded over the ranks must reproduce the single-GPU step over the concatenated batch.

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/manual/mgpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import exemplar_vae_b200 as E  # noqa: E402
from exemplar_vae_b200 import distributed as D  # noqa: E402
from oracle import exvae_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T, N, B = 3000, 1001, 64
    args = O.make_args(model_name=sys.argv[1] if len(sys.argv) > 1 else "vae", number_components=N,
                       training_set_size=T, device="cuda")
    p = O.init_params(args, seed=3)
    data = O.synthetic_dataset(T)
    dataset = torch.utils.data.TensorDataset(data, torch.arange(T).view(-1, 1), torch.zeros(T))
    gen = torch.Generator().manual_seed(11)                 # identical draws on every rank
    bidx = torch.randperm(T, generator=gen)[:B * world]
    x = torch.bernoulli(data[bidx], generator=gen)
    ex_idx = torch.randint(0, T, (N,), generator=gen); ex_idx[:8] = bidx[:8]
    n_eps = 1 if args.model_name == "vae" else 2
    eps = [torch.randn(B * world, 40, generator=gen) for _ in range(n_eps)]

    def fresh():
        m = E.importing_model(args)(args).cuda()
        m.load_state_dict({k: v.detach().clone() for k, v in p.items()})
        m.train()
        return m

    # sharded run: local rows [rank*B, (rank+1)*B)
    sl = slice(rank * B, (rank + 1) * B)
    m = fresh()
    D.shard_bank(m, None, dist.group.WORLD)
    m.rng_override = {"eps": [e[sl].cuda() for e in eps], "exemplar_indices": ex_idx.cuda()}
    loss, RE, KL = m.calculate_loss((x[sl].cuda(), bidx[sl].view(-1, 1).cuda()), 0.7, average=True, dataset=dataset)
    loss.backward()
    m.grad_sync()
    l3 = torch.stack((loss.detach(), RE.detach(), KL.detach()))
    dist.all_reduce(l3); l3 /= world
    ok = True
    if rank == 0:
        ref = fresh()
        ref.rng_override = {"eps": [e.cuda() for e in eps], "exemplar_indices": ex_idx.cuda()}
        l, r, k = ref.calculate_loss((x.cuda(), bidx.view(-1, 1).cuda()), 0.7, average=True, dataset=dataset)
        l.backward()
        r3 = torch.stack((l.detach(), r.detach(), k.detach()))
        ok &= bool(torch.allclose(l3, r3, rtol=1e-5))
        worst = 0.0
        for (n1, a), (n2, b) in zip(m.named_parameters(), ref.named_parameters()):
            scale = b.grad.abs().max().item() + 1e-12
            err = (a.grad - b.grad).abs().max().item() / scale
            worst = max(worst, err)
        ok &= worst < 2e-4
        print(f"mgpu_check world={world} model={args.model_name}: loss sharded={l3.tolist()} single={r3.tolist()} "
              f"worst_rel_grad_err={worst:.2e} -> {'OK' if ok else 'FAIL'}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
