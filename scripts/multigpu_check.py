"""Multi-GPU parity check (launch with torchrun, one rank per GPU):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multigpu_check.py
Every rank runs the full step on its brick (local cells + foreign halo, NCCL
halo exchanges inside swiftgpu_run_step) and compares its LOCAL particles with
the single-rank oracle run on the whole box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import util
from swift_b200 import abi, host
from swift_b200.engine import SwiftGPU, nccl_unique_id

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ok = True
    for scheme in ("minimal", "gadget2", "sphenix"):
        ic = host.jittered_box(24, abi.SCHEMES[scheme], jitter=0.2, h_scatter=0.1, seed=13)
        c = util.make_case(scheme, ic, (4, 4, 4), rank_grid=GRIDS[world], rank=rank)
        c.cfg.device = local_rank
        sub, parts, sel, is_local = host.extract_rank(c.tree, c.parts, c.layout, rank)
        # scramble the foreign copies' h and v: the xv exchange must restore them
        scr = parts.copy()
        hf = host.field(scr, c.layout, "h"); vf = host.field(scr, c.layout, "v")
        hf[~is_local] *= 1.5; vf[~is_local] = 7.0
        g = SwiftGPU(c.cfg)
        g.upload_cells(sub.cells, sub.top)
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        g.halo_setup(idt.cpu().numpy().tobytes())
        g.upload_parts(scr)
        g.set_step(c.step)
        g.run_step(abi.PHASE_ALL)
        got = g.download_parts()
        nd, ng, nf = g.download_counts()
        # oracle: the whole box on one rank
        c1 = util.make_case(scheme, ic, (4, 4, 4))
        o, kind = util.run_oracle(c1)
        # proxies arrive (and stay) in the SENDER's device order, so only a rank's local rows of its
        # AoS array are meaningful: assemble the whole box from every rank's local rows
        size = c.layout.size
        full = torch.zeros(c.n * size, dtype=torch.uint8, device="cuda").reshape(-1, size)
        rows = torch.from_numpy(got.reshape(-1, size)[is_local]).cuda()
        full[torch.from_numpy(sel[is_local]).cuda()] = rows
        full32 = full.to(torch.int32)
        dist.all_reduce(full32, op=dist.ReduceOp.SUM)
        got = full32.to(torch.uint8).cpu().numpy().reshape(-1)[np.repeat(sel, size) * size + np.tile(np.arange(size), sel.shape[0])]
        want = o.parts().reshape(-1, c.layout.size)[sel].reshape(-1)
        rep = util.parity_report(got, want, c.layout, scheme, only=is_local)
        p1 = util.run_port(c1)
        pnd, png, pnf = p1.counts()
        mism = int((nf[is_local] != pnf[sel][is_local]).sum())
        print(f"[rank {rank}/{world}] {scheme} vs {kind}: n_local={int(is_local.sum())} n_foreign={int((~is_local).sum())} "
              f"force-count mismatches={mism} {rep}", flush=True)
        try:
            util.assert_parity(rep)
            assert mism <= max(2, int(2e-3 * is_local.sum()))
        except AssertionError as e:
            ok = False
            print(f"[rank {rank}] FAIL {scheme}: {e}", flush=True)
        g.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
