"""Worker of tests/test_zz_nccl_ranks.py: one process per GPU (torchrun), the library's own NCCL binding behind the all-reduce
hook.  Every rank holds one position range of a small genome; rank 0 also runs the whole genome in one context and the
sharded results must equal it bit for bit: loci and signatures concatenated in rank order, FDR tables on every rank, the
transposons (straddling the cuts, spanning several ranges, nested) on every rank -- also when one rank's signature store is
too small and every rank has to repeat the pass.  Exit code 0 = identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pingpongpro_b200 import api, dist as pdist, host  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def run(lengths, reads, iv, mode, rank, world, local_rank, cuts, pair_capacity=0):
    with api.Context(lengths, mode=mode, device=local_rank, range_begin=cuts[rank], range_end=cuts[rank + 1], pair_capacity=pair_capacity) as ctx:
        pdist.attach_nccl(ctx)
        for c, r in enumerate(reads):
            if len(r):
                ctx.count_reads(c, r)  # broadcast: every shard keeps what it owns (+ its halo)
        ctx.finish()
        ctx.map_heights_to_scores(want_table=False, want_max_height=False)
        ctx.count_stacks_by_group(want_table=False, want_count=False)
        sc = ctx.calculate_fdrs(0)
        tp, order = ctx.find_suppressed_transposons(iv)
        return sc, ctx.signatures(), tp, order


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    failures = []
    for genome, n_reads, seed, mode in (("tiny", 150_000, 21, "weighted"), ("dm6", 400_000, 3, "unique")):
        m = host.SynthModel(genome, n_reads=n_reads, seed=seed)
        reads = m.generate(mode=mode)
        total = sum(m.lengths)
        cuts = [total * k // world for k in range(world + 1)]
        iv3 = m.intervals()
        rows = [(int(c), 0, int(a), int(b) - 1) for c, a, b in iv3[:2000]]
        for c, L in enumerate(m.lengths):
            rows += [(c, 1, 0, L - 1), (c, 0, L // 3, 2 * L // 3), (c, 1, L // 3 + 5, L // 3 + 400)]
        iv = np.zeros(len(rows), api.INTERVAL_DTYPE)
        for k, row in enumerate(rows):
            iv[k] = row
        for caps in (None, [64 if r == world - 1 else 0 for r in range(world)]):
            sc, sigs, tp, order = run(m.lengths, reads, iv, mode, rank, world, local_rank, cuts, pair_capacity=(caps[rank] if caps else 0))
            gathered = [None] * world
            tdist.all_gather_object(gathered, (sc.loci.tobytes(), sigs.tobytes(), sc.n_bins, bits(sc.fdr_table).tobytes(), tp.tobytes(), order.tobytes()))
            if rank == 0:
                with api.Context(m.lengths, mode=mode, device=local_rank) as one:
                    for c, r in enumerate(reads):
                        if len(r):
                            one.count_reads(c, r)
                    one.finish()
                    one.map_heights_to_scores(False)
                    one.count_stacks_by_group(False)
                    ref = one.calculate_fdrs(0)
                    ref_sigs = one.signatures()
                    ref_tp, ref_order = one.find_suppressed_transposons(iv)
                tag = f"{genome}/{mode}/caps={caps}"
                if b"".join(g[0] for g in gathered) != ref.loci.tobytes():
                    failures.append(tag + ": loci differ")
                if b"".join(g[1] for g in gathered) != ref_sigs.tobytes():
                    failures.append(tag + ": signatures differ")
                for r, g in enumerate(gathered):
                    if g[2] != ref.n_bins or g[3] != bits(ref.fdr_table).tobytes():
                        failures.append(f"{tag}: FDR table of rank {r} differs")
                    got = np.frombuffer(g[4], api.TP_RESULT_DTYPE)
                    same = (np.array_equal(bits(got["histogram"]), bits(ref_tp["histogram"])) and np.array_equal(bits(got["reads_plus"]), bits(ref_tp["reads_plus"]))
                            and np.array_equal(got["p"].view(np.uint64), ref_tp["p"].view(np.uint64)) and np.array_equal(bits(got["q_value"]), bits(ref_tp["q_value"]))
                            and g[5] == ref_order.tobytes())
                    if not same:
                        failures.append(f"{tag}: transposons of rank {r} differ")
                print(f"[nccl_worker] {tag}: {len(ref.loci)} loci, {len(ref_sigs)} signatures, {len(iv)} transposons on {world} ranks -- "
                      f"{'identical' if not failures else failures}", flush=True)
    flag = torch.tensor([len(failures)], device="cuda")
    tdist.broadcast(flag, 0)
    tdist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
