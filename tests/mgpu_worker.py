"""Multi-GPU parity worker (launched by torchrun, one rank per GPU): the row-partitioned engine (local Aprod, the
partial A_p'u_p exchanged over NVLink peer memory -- or by one NCCL all-reduce -- every iteration) against the serial
CPU oracle on the same problem.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_worker.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as td


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    import lsqr_b200
    from lsqr_b200 import dist, synth, synth_device
    from oracle import oracle as O      # the checker

    # (workload, scale, tolerance, se wanted, exchange over peer memory (else one NCCL all-reduce), forced blocked layouts)
    cases = (("C2", 20, 1e-10, False, True, False), ("C3", 100, 1e-8, True, True, False), ("C4", 200, 1e-10, False, True, False),
             ("C5", 200, 1e-10, False, True, True), ("C5", 200, 1e-10, True, False, True), ("C3", 100, 1e-8, True, False, False))
    if os.environ.get("MGPU_CASES"):
        cases = tuple(cases[int(i)] for i in os.environ["MGPU_CASES"].split(","))
    for name, scale, tol, want_se, peer, blocked in cases:
        os.environ["LSQR_B200_PEER_EXCHANGE"] = "1" if peer else "0"
        # the layouts of the full-size multi-GPU runs: column-blocked A and row-blocked A' (3 blocks each per rank)
        cfg = synth.scaled(name, scale)
        m, n = cfg["m"], cfg["n"]
        row0, row1 = dist.row_block(m, world, rank)
        for key, val in (("LSQR_B200_VBLOCK_COLS", str(n // 3 + 1)), ("LSQR_B200_UBLOCK_ROWS", str((row1 - row0) // 3 + 1))):
            if blocked:
                os.environ[key] = val
            else:
                os.environ.pop(key, None)
        irow, icol, a = synth_device.coo_block(cfg["kind"], cfg["seed"], m, n, cfg["k"], row0, row1 - row0, dev)
        uid = dist.exchange_unique_id(world, rank)
        s = lsqr_b200.LsqrSolverEz().initialize(row1 - row0, n, a, irow, icol, atol=tol, btol=tol, conlim=1e8, itnlim=4000,
                                                stream=torch.cuda.current_stream().cuda_stream,
                                                world_size=world, rank=rank, nccl_unique_id=uid, m_global=m)
        # the whole problem on the host for the oracle (every rank builds it: small)
        I, J, A = synth.coo_block(cfg["kind"], cfg["seed"], m, n, cfg["k"])
        b = synth.rhs_block(I, J, A, m, synth.x_true(cfg["seed"], n), cfg["seed"])
        plan = s.plan(False)
        if peer:
            assert plan["peer_exchange"] == 1, "the peer-memory exchange could not be set up (IPC mapping failed?)"
        r = s.solve(np.ascontiguousarray(b[row0:row1]), cfg["damp"], want_se=want_se)
        r2 = s.solve(np.ascontiguousarray(b[row0:row1]), cfg["damp"], want_se=want_se)      # reusable and reproducible
        assert r2.itn == r.itn and np.array_equal(np.asarray(r2.x), np.asarray(r.x))
        # every rank must hold the identical solution and scalars
        xs = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(world)]
        td.all_gather(xs, torch.from_numpy(np.asarray(r.x)).to(dev))
        assert all(torch.equal(xs[0], t) for t in xs), "ranks disagree on x"
        if rank == 0:
            ref = O.SolverEz(m, n, A, I, J, atol=tol, btol=tol, conlim=1e8, itnlim=4000).solve(b, cfg["damp"], wantse=want_se)
            rel = np.linalg.norm(np.asarray(r.x) - ref.x) / np.linalg.norm(ref.x)
            assert r.istop == ref.istop, (name, r.istop, ref.istop)
            assert abs(r.itn - ref.itn) <= 2, (name, r.itn, ref.itn)
            assert rel <= 1e-10, (name, rel)
            assert abs(r.rnorm - ref.rnorm) <= 1e-10 * ref.rnorm, (name, r.rnorm, ref.rnorm)
            if want_se:
                rse = np.linalg.norm(np.asarray(r.se) - ref.se) / np.linalg.norm(ref.se)
                assert rse <= 1e-8, (name, rse)
            print(f"MGPU_OK {name}/{scale} peer={int(bool(plan['peer_exchange']))} blocked={blocked} blocks={s.blocks(False)[0]}/{s.blocks(True)[0]} "
                  f"world={world} istop={r.istop} itn={r.itn} (oracle {ref.itn}) rel_x={rel:.2e}", flush=True)
        del s
        td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
