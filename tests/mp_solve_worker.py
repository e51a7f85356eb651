"""torchrun worker of the multi-process tests: every rank solves the same generated system on
its own device (backend nccl), or -- backend gloo, no GPU -- only exercises the wiring with a
stand-in context.  Rank 0 writes a JSON summary."""
import argparse
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeCtx:
    """Stand-in for the CPU wiring test: records what the wiring delivered."""

    def __init__(self, rank):
        self.rank, self.uid, self.blobs = rank, None, None

    def comm_init(self, uid):
        self.uid = uid

    def exchange_export(self):
        return bytes([self.rank]) * 128

    def exchange_import(self, blobs):
        self.blobs = blobs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1536)
    ap.add_argument("--max-iter", type=int, default=150)
    ap.add_argument("--out", required=True)
    ap.add_argument("--backend", default="nccl")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wiring = importlib.import_module("conjugate-gradient_b200.wiring")
    if a.backend == "gloo":
        dist.init_process_group("gloo")
        ctx = FakeCtx(rank)
        wiring.wire(ctx, rank, world, dist, lambda: b"U" * 128)
        first, rows = wiring.shard_of(a.n, rank, world)
        rec = {"rank": rank, "uid": ctx.uid.decode(), "blob_ranks": [b[0] for b in ctx.blobs],
               "first": first, "rows": rows}
        allrec = [None] * world
        dist.all_gather_object(allrec, rec)
        if rank == 0:
            json.dump(allrec, open(a.out, "w"))
        dist.destroy_process_group()
        return 0
    cgb = importlib.import_module("conjugate-gradient_b200")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.n
    ctx = cgb.Context(n, rank, world, local)
    wiring.wire(ctx, rank, world, dist, cgb.unique_id)
    ctx.generate_lap2d()
    ctx.set_rhs(cgb.init_source_term(n))
    summary = {"nblk": ctx.layout().nblk}
    for mode, opt, sched in (("nccl", 0, 0), ("fused", 1, 0), ("persistent", 1, 1)):
        ctx.set_option("exchange", opt)
        ctx.set_option("schedule", sched)
        assert ctx.get_option("schedule_in_use") == sched
        x = np.zeros(n)
        info, hist = ctx.solve(x, max_iter=a.max_iter, tol=1e-10, history=True)
        digest = hashlib.sha256(x.tobytes() + hist.tobytes()).hexdigest()
        digests = [None] * world
        dist.all_gather_object(digests, digest)
        summary[mode] = {"k": int(info.k), "hist": hist.tolist(), "x": x.tolist(),
                         "ranks_agree": len(set(digests)) == 1}
    ctx.close()
    if rank == 0:
        json.dump(summary, open(a.out, "w"))
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
