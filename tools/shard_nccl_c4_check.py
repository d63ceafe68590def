"""N-GPU check of the sharded tiecov path (csrc/shard.cu) under torchrun: one coordinate-sorted stream, split in stream
order at arbitrary record indices (inside bundles), tc_shard_coverage (NCCL halo exchange) + tc_shard_gather on every rank;
rank 0 compares the gathered rows with the single-GPU tc_coverage_stream of the whole stream and with the oracle.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_nccl_c4_check.py [records]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tiebrush_b200 import api, synth


def slice_cols(cols, a, b):
    """records [a,b) of a torch column dict as a device-resident segment with its own CIGAR arena"""
    off = cols["cig_off"].to(torch.int64) & 0xFFFFFFFF
    c0, c1 = int(off[a]), int(off[b])
    out = {k: cols[k][a:b].contiguous() for k in ("tid", "pos", "yc", "strand")}
    out["cig_off"] = (off[a:b + 1] - c0).to(torch.int32).contiguous()
    out["cigar"] = cols["cigar"][c0:c1].contiguous()
    out["n_cig"] = c1 - c0
    return out


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for case, (chroms, n_tx, seed) in enumerate([(3, 60, 11), (1, 8, 12), (5, 300, 13)]):
        cols = synth.coverage_stream(n, seed=seed, n_tx=n_tx, chroms=chroms, device=dev)    # the same stream on every rank
        host = synth.to_host(cols)
        # cuts in stream order at awkward places: inside bundles, one rank almost empty
        rng = np.random.default_rng(100 + case)
        cuts = np.sort(rng.integers(1, n - 1, world - 1)) if world > 1 else np.zeros(0, np.int64)
        if case == 1 and world > 2:
            cuts[1] = cuts[0] + 3          # a rank of three records (inside one bundle with n_tx = 8: no head of its own)
        bounds = [0] + [int(c) for c in cuts] + [n]
        a, b = bounds[rank], bounds[rank + 1]
        seg = slice_cols(cols, a, b)
        with api.Context(device=local) as ctx:
            idt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            ctx.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
            cap = 2 * int(host["cig_off"][-1]) + 64
            loc = ctx.shard_coverage([seg], 50000, api.cov_out_buffers(cap, cap, device=dev))
            g = ctx.shard_gather(loc, api.cov_out_buffers(cap, cap, device=dev) if rank == 0 else None, world)
            stats = loc["stats"]
            print(f"case {case} rank {rank}: records [{a},{b}) runs {loc['n_runs']} juncs {loc['n_juncs']} windows {loc['windows']} {stats} gather_bytes {g['gather_bytes']}", flush=True)
            # the same with the gather folded into the window loop (regions on rank 0)
            ov = ctx.shard_coverage_gather([seg], 50000, api.cov_out_buffers(cap, cap, device=dev), api.cov_out_buffers(cap * world, cap * world, device=dev) if rank == 0 else None,
                                           cap * world, cap * world, world)
            print(f"case {case} rank {rank}: overlapped gather: rounds {ov['stats']['gather_rounds']} bytes {ov['stats']['gather_bytes']} regions {ov['regions']}", flush=True)
            if rank == 0:
                one = ctx.coverage_stream(cols, 50000, api.cov_out_buffers(cap, cap, device=dev))
                for name, x, y in zip(("r_tid", "r_start", "r_end", "r_val", "j_tid", "j_start", "j_end", "j_strand", "j_val"),
                                      ov["gathered_runs"]() + ov["gathered_juncs"](), one["runs"] + one["juncs"]):
                    if not np.array_equal(x.cpu().numpy(), y.cpu().numpy()):
                        ok = False
                        print(f"case {case}: overlapped gather: {name} differs: {len(x)} vs {len(y)}", flush=True)
                from oracle import oracle
                exp = oracle.coverage(host)
                for name, x, y, z in zip(("r_tid", "r_start", "r_end", "r_val", "j_tid", "j_start", "j_end", "j_strand", "j_val"),
                                         g["runs"] + g["juncs"], one["runs"] + one["juncs"], exp["runs"] + exp["juncs"]):
                    x, y = x.cpu().numpy(), y.cpu().numpy()
                    if not (np.array_equal(x, y) and np.array_equal(y, z)):
                        ok = False
                        print(f"case {case}: {name} differs: sharded {len(x)} single {len(y)} oracle {len(z)}", flush=True)
                print(f"case {case}: sharded x{world} == single GPU == oracle: {ok}; junc_base {g['junc_base']}", flush=True)
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("SHARD_NCCL_C4_CHECK OK", flush=True)


if __name__ == "__main__":
    main()
