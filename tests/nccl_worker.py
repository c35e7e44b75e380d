"""Worker of tests/test_gpu_multi.py (one rank per GPU, launched by torch.distributed.run): the in-library NCCL
exchange (dist.NcclFit -> pmf_fit) or the host-driven one (dist.ShardedFit) on a row shard; rank 0 also runs the
single-handle fit of the whole problem and writes both to a JSON file."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import pathmatfac_b200 as P
from pathmatfac_b200 import _lib
from pathmatfac_b200.dist import NcclFit, ShardedFit, shard_rows
from tests.helpers import make_pair


def main():
    out_path, mode, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    M, n_feat, K = int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    views = {"mutation": ("bernoulli", n_feat // 4), "methylation": ("normal", n_feat // 2), "counts": ("poisson", n_feat // 4)}
    model, om, D = make_pair(M, views, K=K, seed=77, missing=0.25, lambda_X_l2=1.0, batch_views=["methylation"], n_batches=4)
    kern = {"ffma": _lib.KERNEL_FFMA, "tc": _lib.KERNEL_TC, "auto": _lib.KERNEL_AUTO}[kernel]
    epochs = 8
    common = dict(lr=0.25, update_X=1, update_Y=1, update_col_layers=1, kernel=kern, rel_tol=0.0, abs_tol=0.0, check_every=3)
    full = None
    if rank == 0:          # the concatenated problem on one handle
        import copy
        m1 = copy.deepcopy(model)
        e1 = P.Engine(m1, device=local)
        e1.reset_opt_state(1e-8)
        h1 = e1.fit(e1.make_opts(epoch=1, max_epochs=epochs, **common))
        e1.pull_params()
        e1.close()
        full = dict(loss=h1["loss"], Y=np.asarray(m1.matfac.Y).tolist(), X=np.asarray(m1.matfac.X).tolist(),
                    theta=np.asarray(m1.matfac.col_transform.unwrapped(3).theta.values[0]).tolist(), term=h1["term_code"])
    rows = shard_rows(M, rank, world)
    eng = P.Engine(model, device=local, rows=rows)
    eng.reset_opt_state(1e-8)
    drv = NcclFit(eng) if mode == "nccl" else ShardedFit(eng)
    h = drv.fit(eng.make_opts(epoch=1, max_epochs=epochs, **common))
    eng.pull_params()
    Y = torch.from_numpy(np.ascontiguousarray(model.matfac.Y)).cuda()
    Ys = [torch.empty_like(Y) for _ in range(world)]
    dist.all_gather(Ys, Y)
    Xloc = torch.zeros((K, M), device="cuda")
    Xloc[:, rows.start:rows.stop] = torch.from_numpy(np.ascontiguousarray(model.matfac.X[:, rows.start:rows.stop])).cuda()
    dist.all_reduce(Xloc)
    th = model.matfac.col_transform.unwrapped(3).theta.values[0]
    if mode == "nccl":
        drv.close()
    eng.close()
    if rank == 0:
        res = dict(world=world, mode=mode, loss=h["loss"], term=h["term_code"], full=full, launches=h["kernel_launches"],
                   Y_equal_across_ranks=bool(all(torch.equal(Ys[0], y) for y in Ys)),
                   Y=Ys[0].cpu().numpy().tolist(), X=Xloc.cpu().numpy().tolist(), theta=np.asarray(th).tolist())
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
