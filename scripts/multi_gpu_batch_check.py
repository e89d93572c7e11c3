"""Run under torchrun: a batch of independent conic problems split across the ranks (batch split; no
collective on the data path) must give, problem for problem, the bits a single handle holding the whole batch gives.
  --same-device   every rank uses cuda:0 (gloo rendezvous): runs on a box with one GPU.
Exit code 0 = every rank agrees."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fos_b200 as fos  # noqa: E402
from fos_b200 import parallel, problems  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    one_dev = "--same-device" in sys.argv
    if one_dev:
        local = 0
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    B = 37   # not a multiple of the rank count: ragged shards
    plist = [problems.nnls_conic(24, 30, seed=100 + j, scale=0.3) for j in range(B)]
    A = np.stack([np.asarray(P.A.todense()) for P in plist])
    b = np.stack([P.b for P in plist])
    c = np.stack([P.c for P in plist])
    ok = True
    for name, alg in (("FISTA", fos.FISTA()), ("Dykstra", fos.Dykstra())):
        p0, cnt = parallel.batch_shard(B, rank, world)
        H = fos.Handle(local)
        H.load_conic_batch(A[p0:p0 + cnt], b[p0:p0 + cnt], c[p0:p0 + cnt], plist[0].constr_cones, plist[0].var_cones)
        H.set_algorithm(alg)
        done, st, recs, guess = H.solve_batch(300, 25, 1e-3)
        mine = torch.zeros((B, guess.shape[1] + 2), dtype=torch.float64)
        mine[p0:p0 + cnt, :-2] = torch.from_numpy(guess)
        mine[p0:p0 + cnt, -2] = torch.from_numpy(done.astype(np.float64))
        mine[p0:p0 + cnt, -1] = torch.from_numpy(st.astype(np.float64))
        dist.all_reduce(mine)   # shards are disjoint: the sum is the gathered result (host side, results only)
        if rank == 0:
            H1 = fos.Handle(local)
            H1.load_conic_batch(A, b, c, plist[0].constr_cones, plist[0].var_cones)
            H1.set_algorithm(alg)
            d1, s1, r1, g1 = H1.solve_batch(300, 25, 1e-3)
            same = (np.array_equal(mine[:, :-2].numpy(), g1) and np.array_equal(mine[:, -2].numpy(), d1.astype(np.float64))
                    and np.array_equal(mine[:, -1].numpy(), s1.astype(np.float64)))
            print(f"{name}: {world} shards vs one batch of {B}: {'bitwise identical' if same else 'DIFFERENT'}; "
                  f"iterations {sorted(set(d1.tolist()))[:4]}...", flush=True)
            ok = ok and same
        del H
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
