"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md §8e.

dense   : the sequence is cut into window ranges; each rank reads its bytes plus
          a (k-1)-byte halo and counts only windows STARTING in its range, so no
          carry is exchanged; the uint32[4^k] tables are summed with one reduce
          (int32 views: two's-complement add == uint32 add mod 2^32).
per-seq : the reference-shaped table int32[4^k][num_seqs] (kernels.h:142): sequences are
          independent, so each rank counts a range of them (cut at sequence starts, equal
          bytes per rank) and the columns are all-gathered; every rank ends with the whole
          table, which is what the distance step (kernels.h:85-109) reads.
sparse  : reads are cut by read index; each rank counts locally, buckets its
          (code, count) pairs by owner = mix64(code) % world, exchanges buckets
          with one all-to-all, and the owner merges what it received: every rank
          ends with a disjoint, sorted key range of the global result.

The functions take the local "engine" as callables so the same code runs on the
GPU (kmerb200.Context) and in the gloo tests (the oracle stands in on the CPU).
"""
import numpy as np

from . import shard_reads, shard_seqs, shard_windows  # noqa: F401  (re-exported)


def _dist():
    import torch.distributed as dist
    return dist


def reduce_table(table, dst=0):
    """Sum uint32 tables (held as int32 torch tensors) onto rank `dst`."""
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(table, dst=dst, op=dist.ReduceOp.SUM)
    return table


def all_reduce_table(table):
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM)
    return table


def exchange_by_owner(keys, counts, sizes):
    """keys (int64) / counts (int32) torch tensors laid out owner by owner with
    `sizes[o]` entries for owner o.  Returns what this rank owns: (keys, counts)
    concatenated over senders."""
    import torch
    dist = _dist()
    world = dist.get_world_size()
    dev = keys.device
    send = torch.tensor([int(s) for s in sizes], dtype=torch.int64, device=dev)
    recv = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send)
    in_splits = [int(x) for x in send.cpu().tolist()]
    out_splits = [int(x) for x in recv.cpu().tolist()]
    rkeys = torch.empty(sum(out_splits), dtype=keys.dtype, device=dev)
    rcounts = torch.empty(sum(out_splits), dtype=counts.dtype, device=dev)
    dist.all_to_all_single(rkeys, keys[: sum(in_splits)].contiguous(), out_splits, in_splits)
    dist.all_to_all_single(rcounts, counts[: sum(in_splits)].contiguous(), out_splits, in_splits)
    return rkeys, rcounts


def count_dense_sharded(count_range, make_shard, nbytes, k, table, rank, world, dst=0):
    """count_range(shard, shard_nbytes, win_begin, win_end, table) adds into `table`;
    make_shard(byte_begin, byte_end) returns this rank's bytes.  The reduced table
    lands on rank `dst`."""
    b, e, bb, be = shard_windows(nbytes, k, rank, world)
    if e > b:
        shard = make_shard(bb, be)
        count_range(shard, be - bb, 0, e - b, table)
    return reduce_table(table, dst)


def count_per_seq_sharded(count_seqs, offsets, rank, world):
    """SURVEY §8e, per-sequence mode.  count_seqs(s0, s1) returns this rank's columns, an int32 tensor
    [4^k, s1 - s0] (k-mer-major like the reference's `sums`, kernels.h:142) for the sequences [s0, s1) — the
    caller uploads only those sequences' bytes (offsets[s0] .. offsets[s1]).  Returns the whole table
    [4^k, num_seqs] on every rank: the columns are all-gathered (padded to the widest shard, one collective)."""
    import torch
    dist = _dist()
    n = len(offsets) - 1
    cuts = [shard_seqs(offsets, r, world) for r in range(world)]
    s0, s1 = cuts[rank]
    local = count_seqs(s0, s1)
    if world == 1 or not dist.is_initialized():
        return local
    nk = local.shape[0]
    widest = max(max(c[1] - c[0] for c in cuts), 1)
    padded = torch.zeros((nk, widest), dtype=local.dtype, device=local.device)
    if s1 > s0:
        padded[:, : s1 - s0] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = torch.cat([parts[r][:, : cuts[r][1] - cuts[r][0]] for r in range(world)], dim=1).contiguous()
    assert out.shape == (nk, n)
    return out


def count_sparse_radix_sharded(engine, reads, nbytes, k, table_full=()):
    """Range-sharded radix path: every rank scatters its reads into the 1024 level-1 partitions
    (top 10 bits of the code), ONE equal-split all-to-all moves each rank's partition range to it
    (the slabs are partition-major, so a range is one contiguous block), and every rank counts the
    partitions it owns: rank r ends with the sorted k-mers of code range r, nothing is merged.

    `engine` provides radix_plan / radix_scatter / radix_count (kmerb200.Context on the GPU; the
    CPU tests pass an adapter over the emulator build).  Returns None when any rank overflowed
    (skewed input): the caller then takes the hash-sharded path.  `table_full` = exception
    types that mean "overflow" for this engine.  ANY exception on one rank (out of memory, a bad
    argument) is voted on like an overflow before it is raised again, so that no rank is left
    waiting in a collective its peer never enters."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(), dist.get_rank()

    def vote(state, dev):
        """state: 2 = ok, 1 = overflow, 0 = error; returns the minimum over the ranks"""
        t = torch.tensor([state], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def stage(f, dev):
        """run f() on this rank, agree on the outcome with every rank"""
        res, err, state = None, None, 2
        try:
            res = f()
        except table_full:
            state = 1
        except Exception as ex:  # noqa: BLE001 — voted on, then raised again below
            err, state = ex, 0
        worst = vote(state, dev)
        if err is not None:
            raise err
        if worst == 0:
            raise RuntimeError("sharded sparse radix: another rank failed (see its error)")
        return res if worst == 2 else None

    dev = reads.device if hasattr(reads, "device") else "cpu"
    t = torch.tensor([max(nbytes - k + 1, 0)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # Large inputs run in 2^round_bits rounds (round r = the windows whose top code bits are r; the plan sizes them so
    # that a leaf fits shared memory and the slabs fit the device): every round has its own scatter, all-to-all and
    # count and reuses the buffers; rank r's rounds are ascending code ranges, appended to one result.  An overflow on any
    # rank (leaves denser than planned) is retried twice with one more round bit before the caller's fallback.
    min_bits, bufs, recv = 0, None, None
    for attempt in range(3):
        try:
            plan = stage(lambda: engine.radix_plan(int(t.item()), k, world, min_bits), dev)
        except Exception:
            if attempt == 0:
                raise
            return None  # no code bits left for another round
        if plan is None:
            return None
        if bufs is not None and (bufs[0].numel() != plan.slab_bytes or bufs[1].numel() * 4 != plan.counts_bytes):
            bufs = recv = None  # the new plan has other buffer sizes
        rounds = 1 << getattr(plan, "round_bits", 0)
        acc, ok = None, True  # one result that the rounds append to (kc_sparse_radix_count_round_append): no pieces, no concat
        for rnd in range(rounds):
            sc = stage(lambda: engine.radix_scatter(reads, nbytes, plan, rnd, bufs), dev)
            if sc is None:
                ok = False
                break
            bufs = sc
            if recv is None:
                recv = stage(lambda: (torch.empty_like(bufs[0]), torch.empty_like(bufs[1])), dev)
            dist.all_to_all_single(recv[0], bufs[0])  # equal splits: block o = partitions [o, o+1) * parts_per_rank
            dist.all_to_all_single(recv[1], bufs[1])
            res = stage(lambda: engine.radix_count_append(plan, recv[0], recv[1], world, rank * plan.parts_per_rank, plan.parts_per_rank,
                                                          rnd, acc), dev)
            if res is None:
                ok = False
                break
            acc = res
        if ok:
            return engine.finish(acc) if hasattr(engine, "finish") else acc
        if acc is not None:
            if hasattr(acc, "close"):
                acc.close()
            elif hasattr(engine, "finish"):
                engine.finish(acc)
        min_bits = getattr(plan, "round_bits", 0) + 1
        if min_bits > 8:
            break
    return None


def count_sparse_sharded_gpu(ctx, d_reads, nbytes, k, algo=0):
    """GPU path of the sharded sparse count: `d_reads` holds THIS rank's reads.  Returns a
    kmerb200.Sparse with the keys this rank owns: a code RANGE with SPARSE_RADIX (sorted
    across ranks), the codes with mix64(code) % world == rank otherwise."""
    import torch
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return ctx.count_sparse(d_reads, nbytes, k, algo)
    from . import KC_ERR_TABLE_FULL, SPARSE_AUTO, SPARSE_HASH, SPARSE_NO_FALLBACK, SPARSE_RADIX, SPARSE_UNSORTED, KmerError
    if (algo & 0xFF) == SPARSE_AUTO:  # as kc_count_sparse: radix wherever it exists
        algo = (algo & ~0xFF) | (SPARSE_RADIX if 2 * k > 20 else SPARSE_HASH)
    if (algo & 0xFF) == SPARSE_RADIX and 1024 % world != 0:
        if algo & SPARSE_NO_FALLBACK:
            raise KmerError(-1, "range-sharded radix needs a world size that divides 1024 (got %d)" % world)
        algo = SPARSE_HASH  # the documented fallback: owner-hash sharding works for any world size
    if (algo & 0xFF) == SPARSE_RADIX:
        class _Full(KmerError):
            pass

        class _Engine:  # turns KC_ERR_TABLE_FULL into its own exception type, passes the rest on
            def __getattr__(self, name):
                f = getattr(ctx, name)

                def call(*a):
                    try:
                        return f(*a)
                    except KmerError as e:
                        if e.code == KC_ERR_TABLE_FULL:
                            raise _Full(e.code, str(e))
                        raise
                return call

        res = count_sparse_radix_sharded(_Engine(), d_reads, nbytes, k, table_full=(_Full,))
        if res is not None:
            return res
        if algo & SPARSE_NO_FALLBACK:
            raise KmerError(KC_ERR_TABLE_FULL, "sharded sparse radix overflowed on some rank (skewed input)")
        algo = SPARSE_HASH
    local = ctx.count_sparse(d_reads, nbytes, k, algo | SPARSE_UNSORTED)  # re-bucketed below: no local sort
    n = len(local)
    dev = "cuda:%d" % ctx.device
    ok = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    oc = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    sizes = ctx.sparse_bucket_by_owner(local.d_keys, local.d_counts, n, world, ok, oc)
    local.close()
    rk, rc = exchange_by_owner(ok, oc, sizes)
    torch.cuda.synchronize()
    return ctx.sparse_merge(rk, rc, rk.numel())


# ---- numpy stand-ins used by the CPU (gloo) tests ---------------------------
def mix64_np(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xC4CEB9FE1A85EC53)
    x ^= x >> np.uint64(33)
    return x


def bucket_by_owner_np(keys, counts, world):
    owner = (mix64_np(keys) % np.uint64(world)).astype(np.int64)
    order = np.argsort(owner, kind="stable")
    sizes = np.bincount(owner, minlength=world)
    return keys[order], counts[order], sizes


def merge_np(keys, counts):
    if keys.size == 0:
        return keys, counts
    order = np.argsort(keys, kind="stable")
    k, c = keys[order], counts[order].astype(np.uint64)
    uniq, start = np.unique(k, return_index=True)
    return uniq, np.add.reduceat(c, start).astype(np.uint32)
