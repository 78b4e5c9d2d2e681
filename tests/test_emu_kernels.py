"""Kernel LOGIC on the CPU: the product's .cu sources compiled against the test-only SIMT
emulator (tests/emu) and compared with the oracle.  Runs where there is no GPU, so every
change to a kernel meets the oracle before it meets a B200; the `-m gpu` tests remain the
parity tests proper.  Each case is a subprocess because the emulator reads its scheduling
seed once per process (KC_EMU_SEED != 0: random fiber scheduling with preemption inside the
shared-memory helpers, which is what exercises the barrier-free staging protocols)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "emu", "emu_harness.py")


@pytest.fixture(scope="session", autouse=True)
def _built():
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "tests", "emu")], check=True)


def run_case(*args, seed=0, sms=4, shift=3, **extra_env):
    env = dict(os.environ, KC_EMU_SEED=str(seed), KC_EMU_SMS=str(sms), KC_EMU_PREEMPT_SHIFT=str(shift), KC_EMU_STATS="1")
    env.update({k: str(v) for k, v in extra_env.items()})
    r = subprocess.run([sys.executable, HARNESS] + [str(a) for a in args], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout + r.stderr


def emu_stats(out):
    """the emulator's rare-path counters ([i]=n ...) printed at exit"""
    line = [ln for ln in out.splitlines() if ln.startswith("simt_emu stats:")][-1]
    return {int(t[1:t.index("]")]): int(t.split("=")[1]) for t in line.split() if t.startswith("[")}


# (k, bytes, algo, input kind, data seed, misalignment of the device pointer)
DENSE = [
    (3, 100_000, 0, "dirty", 1, 3),
    (7, 200_000, 0, "genome", 2, 0),
    (8, 300_000, 1, "genome", 2, 9),
    (8, 4_500_000, 0, "genome", 4, 0),   # 16-bit shared-memory bins + reduce
    (8, 4_300_000, 0, "polyA", 4, 1),    # one bin takes every hit: the 0x4000 spill path
    (12, 400_000, 2, "genome", 2, 0),    # partition path
    (12, 400_000, 2, "dirty", 5, 7),
    (12, 300_000, 2, "skew", 6, 2),      # region overflow -> RED fallback
    (11, 300_000, 2, "genome", 7, 15),
    (10, 300_000, 2, "dirty", 8, 4),
    (9, 300_000, 2, "genome", 5, 1),
    (13, 150_000, 0, "dirty", 3, 0),
]


@pytest.mark.parametrize("case", DENSE, ids=lambda c: "k%d-%s-a%d" % (c[0], c[3], c[2]))
def test_dense_kernels_on_emulator(case):
    run_case("dense", *case, seed=0)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_partition_staging_protocol_random_schedules(seed):
    # random interleavings of the 1024 threads of a CTA inside the staging protocol
    run_case("dense", 12, 300_000, 2, "genome", 10 + seed, seed, seed=seed, sms=2)
    run_case("dense", 12, 200_000, 2, "skew", 20 + seed, 0, seed=seed, sms=1, shift=1)


RADIX, NOFB, UNSORTED = 2, 0x200, 0x100
@pytest.mark.parametrize("seed", [1, 2])
def test_partition_deferred_retry_variant(seed):
    """KC_PART_ABLATE=3: a record that meets a full bin gets a second attempt at the lane's next
    record slot before the RED fallback.  Skewed input under a random schedule makes bins fill
    while their flush is in flight; the counters prove the paths ran."""
    out = run_case("dense", 12, 200_000, 2, "skew", 30 + seed, seed, seed=seed, sms=1, shift=1, KC_PART_ABLATE=3)
    st = emu_stats(out)
    assert st[1] > 100 and st[2] > 0 and st[1] > st[2], st  # deferred, fell back after the retry, some retries succeeded
    run_case("dense", 12, 300_000, 2, "genome", 40 + seed, 0, seed=seed, sms=2, KC_PART_ABLATE=3)
    run_case("dense", 9, 200_000, 2, "dirty", 50 + seed, 3, seed=0, sms=2, KC_PART_ABLATE=3)


def test_partition_paired_count_variant():
    """KC_DENSE_PARTITION_PAIR (algo 5, k = 12): pass 2 counts the 13-mers at offsets 0 and 2 and the
    12-mer at offset 4 of every record (3 increments instead of 5) and folds them into 12-mer bins at
    the flush; a partition whose 16-bit fields wrapped is recounted with 32-bit bins (forced here for
    every odd partition: a real wrap needs 0.7 GB of input)."""
    run_case("dense", 12, 400_000, 5, "genome", 2, 0)
    run_case("dense", 12, 400_000, 5, "dirty", 5, 7, seed=3, sms=2)
    run_case("dense", 12, 300_000, 5, "skew", 6, 2)
    st = emu_stats(run_case("dense", 12, 400_000, 5, "genome", 8, 3, KC_EMU_FORCE_PAIR_RECOUNT=1))
    assert st[7] >= 1024, st   # the odd partitions went through the 32-bit recount


def test_partition_two_increment_count_variant():
    """KC_DENSE_PARTITION_TRIO (algo 6, k = 12): one 14-mer (8-bit fields) + one 13-mer (16-bit fields)
    per record = 2 increments instead of 5; poly-A makes 8-bit fields wrap for real (partition 0 holds
    ~500 identical records): the checksum must catch it and the 32-bit recount must be exact."""
    run_case("dense", 12, 400_000, 6, "genome", 2, 0)
    run_case("dense", 12, 400_000, 6, "dirty", 5, 7, seed=3, sms=2)
    run_case("dense", 12, 300_000, 6, "skew", 6, 2)
    st = emu_stats(run_case("dense", 12, 2_500_000, 6, "polyA", 6, 2))
    assert st[7] == 2048, st   # partition 0, twice (whole table + window sub-range), 1024 threads each


def test_partition_combined_variants():
    """algo 8 / 9: the deferred-retry scatter (4) with the paired (5) / two-increment (6) count — host dispatch only,
    the kernels are the ones of the tests above"""
    for algo in (8, 9):
        st = emu_stats(run_case("dense", 12, 300_000, algo, "skew", 6, 2, seed=3, sms=2, shift=1))
        assert st[1] > 0, st   # the deferred retry ran


def test_partition_wide_record_variant():
    """KC_DENSE_PARTITION_WIDE (algo 7, k = 12): seven windows per record (18 bases; the slab record omits
    the 11 key bits), pass 2 with two 4-bit 14-mer tables + one 16-bit 12-mer table.  Random schedule on
    skewed input: deferred retries, nibble wraps -> 32-bit recount; poly-A; forced recount."""
    run_case("dense", 12, 400_000, 7, "genome", 2, 0)
    run_case("dense", 12, 400_000, 7, "dirty", 5, 7)
    st = emu_stats(run_case("dense", 12, 300_000, 7, "skew", 6, 2, seed=3, sms=2, shift=1))
    assert st[1] > 100 and st[7] >= 1024, st   # deferred retries happened; at least one partition wrapped and was recounted
    st = emu_stats(run_case("dense", 12, 2_500_000, 7, "polyA", 6, 2))
    assert st[7] >= 1024, st


def test_k8_checksum_variant():
    """KC_DENSE_SMEM16C (algo 3): non-returning shared adds, per-CTA checksum, repair of the CTAs
    whose 16-bit fields wrapped.  Uniform input: no CTA is repaired; one-bin inputs: every CTA is."""
    st = emu_stats(run_case("dense", 8, 4_500_000, 3, "genome", 4, 0))
    assert st[6] == 0, st
    st = emu_stats(run_case("dense", 8, 4_400_000, 3, "polyA", 4, 1))
    assert st[6] >= 4, st          # whole table and the window sub-range: all 4 CTAs each time
    st = emu_stats(run_case("dense", 8, 4_400_000, 3, "skew", 4, 7))
    assert st[6] >= 1, st
    run_case("dense", 8, 4_300_000, 3, "dirty", 5, 2, seed=3, sms=2)
    run_case("dense", 8, 100_000, 3, "genome", 5, 2)  # too small for the interior kernel: direct path


SPARSE = [
    (21, 60_000, 0, "reads", 1, 0),
    (31, 60_000, 0, "dirty", 2, 5),
    (15, 50_000, 1, "genome", 3, 0),
    (21, 60_000, 0 | UNSORTED, "readsU", 4, 0),
]


@pytest.mark.parametrize("case", SPARSE, ids=lambda c: "k%d-%s-a%d" % (c[0], c[3], c[2]))
def test_sparse_kernels_on_emulator(case):
    run_case("sparse", *case, seed=0)


# KC_SPARSE_RADIX with the 16 x 16 test shape (KC_SPARSE_RADIX_SHAPE=small): regions fill up,
# leaves hold thousands of records, every record width combination (32/32, 64/32, 64/64 bits)
# and both scanner halos are hit.  NO_FALLBACK: the radix path itself must produce the result.
@pytest.mark.parametrize("k", [13, 17, 18, 19, 21, 27, 28, 31])
def test_sparse_radix_small_shape(k):
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        run_case("sparse", k, 60_000, RADIX | NOFB, "readsU", k, k % 16, seed=0)
        run_case("sparse", k, 50_000, RADIX | NOFB, "dirty", k + 1, 0, seed=0)
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_sparse_radix_leaf_table_all_ones_code():
    """16 x 16 shape, k = 20: the leaf records keep all 32 bits of uint32, so the all-T window's record equals the leaf
    table's EMPTY marker and is counted on the side (s_special); k = 21 / 31: EMPTY is not a record, same inputs"""
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        for k in (20, 21, 31):
            run_case("sparse", k, 30_000, RADIX | NOFB, "runsT", k, 0, seed=0)
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_sparse_radix_leaf_with_too_many_distinct_codes_gets_more_rounds():
    """16 x 16 shape, 64-bit leaf records: 512 distinct codes per leaf at most; 180 000 bases of shallow reads put ~540 in a
    leaf, inside the plan's record rule — the leaf table fills, the count is void and kc_sparse_radix retries with one
    more round bit (stat 11), which halves the leaves"""
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        out = run_case("sparse", 23, 180_000, RADIX | NOFB, "readsU", 3, 0, seed=0)
        assert emu_stats(out)[11] == 1, out
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_sparse_radix_rounds_on_one_gpu_append_to_one_result():
    """four forced rounds (KC_SPARSE_RADIX_RBITS=2): round 0's result sizes the arrays the later rounds append to; with
    'fewA' the last round does not fit behind the others and is concatenated as a piece of its own"""
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        for k in (21, 31):
            run_case("sparse", k, 60_000, RADIX | NOFB, "readsU", k, 0, seed=0, KC_SPARSE_RADIX_RBITS=2)
        out = run_case("sparse", 17, 60_000, RADIX | NOFB, "fewA", 3, 0, seed=0, KC_SPARSE_RADIX_RBITS=2)
        assert emu_stats(out)[14] >= 1, out
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


@pytest.mark.parametrize("rbits,k", [(1, 13), (3, 17), (3, 18), (5, 24), (4, 31), (8, 31)])
def test_sparse_radix_round_filter(rbits, k):
    """sp_scatter_kernel picks the windows of a round with one bit-plane compare per lane and group (all 16 windows at
    once): every round count, both scanner halos, 'dirty' input with invalid bytes everywhere"""
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        n = 25_000 if rbits < 5 else 6_000  # (2^rbits rounds of 16 x 16 partitions each on the emulator)
        run_case("sparse", k, n, RADIX | NOFB, "dirty", 10 * rbits + k, k % 7, seed=0, KC_SPARSE_RADIX_RBITS=rbits)
        if rbits < 8:  # (256 rounds once are enough)
            run_case("sparse", k, n, RADIX | NOFB, "readsU", rbits + k, 0, seed=0, KC_SPARSE_RADIX_RBITS=rbits)
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_sparse_radix_count_round_append():
    """the multi-GPU form of the appended rounds (kc_sparse_radix_count_round_append, one call per round): arrays sized
    from the first round, grown when a later round does not fit ('fewA': stat 14), a single round adopted as it is"""
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        run_case("radix_append", 21, 40_000, "readsU", 4, seed=0, KC_SPARSE_RADIX_RBITS=3)
        out = run_case("radix_append", 17, 60_000, "fewA", 3, seed=0, KC_SPARSE_RADIX_RBITS=2)
        assert emu_stats(out)[14] >= 1, out
        run_case("radix_append", 31, 20_000, "dirty", 5, seed=0)
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_sparse_radix_short_run_list_is_counted_again_with_the_exact_size():
    """the temporary run list is sized from an estimate; when it is too short the leaf kernel still reports the number of
    runs and the count is repeated once with exactly that many entries (stat 13), no new scatter, no hash fallback"""
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        out = run_case("sparse", 21, 60_000, RADIX | NOFB, "readsU", 5, 0, seed=0, KC_SPARSE_RADIX_RUNLIST=2000)
        st = emu_stats(out)
        assert st[13] == 1 and st[11] == 0, out
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sparse_radix_staging_random_schedules(seed):
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        # deep coverage of a tiny genome: a few partitions take most records, so under an
        # adversarial schedule a writer may find both bins full 256 times and give up —
        # then the hash recount must still deliver the exact result (no NO_FALLBACK here)
        run_case("sparse", 21, 80_000, RADIX, "reads", seed, 3, seed=seed, sms=2, shift=1)
        # (16 partitions x 2 bins for 1024 threads is 64x the contention of the shipped shape:
        # with preemption at every second helper call a writer can starve for 256 attempts)
        run_case("sparse", 21, 60_000, RADIX | NOFB, "readsU", 10 + seed, 5, seed=seed, sms=2, shift=3)
        run_case("sparse", 31, 60_000, RADIX | NOFB, "readsU", seed, 0, seed=seed, sms=1, shift=2)
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_sparse_radix_overflow_falls_back_to_hash():
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        run_case("sparse", 21, 60_000, RADIX, "polyA", 1, 0, seed=0)       # one leaf takes everything
        with pytest.raises(AssertionError, match="overflowed"):
            run_case("sparse", 21, 60_000, RADIX | NOFB, "polyA", 1, 0, seed=0)
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_sparse_radix_production_shape():
    # 1024 x 1024 partitions: mostly empty leaves at this size, but the shipped geometry
    # (2 000 reads: three quarters of the 2^20 leaves stay empty and cost no barriers — 40 000 took 66 s here)
    run_case("sparse", 21, 2_000, RADIX | NOFB, "readsU", 9, 0, seed=0, sms=2)
    if os.environ.get("KC_RUN_SLOW"):
        run_case("sparse", 21, 40_000, RADIX | NOFB, "readsU", 9, 0, seed=0, sms=2)


def test_host_entry_points_on_emulator():
    run_case("dense_host", 5, 300_000, 1)
    run_case("dense_host", 12, 200_000, 2)


def test_perseq_distance_and_generators_on_emulator():
    run_case("perseq", 3, 37, 1)     # shared-memory bins per (sequence, tile) segment
    run_case("perseq", 6, 9, 2)
    run_case("perseq", 8, 5, 3)      # global atomics
    run_case("gen", 200_000, 0xB2000003)


@pytest.mark.parametrize("k,nreads,world", [(21, 600, 4), (31, 500, 2), (15, 500, 8)])
def test_sparse_radix_range_sharded_ranks_emulated(k, nreads, world):
    """the multi-GPU radix path with the ranks run one after the other on the emulator (numpy
    slicing stands in for the all-to-all); tests/test_sharding_gloo.py runs the same with real
    processes and gloo"""
    os.environ["KC_SPARSE_RADIX_SHAPE"] = "small"
    try:
        run_case("radix_sharded", k, nreads, world, 7 + world)
    finally:
        del os.environ["KC_SPARSE_RADIX_SHAPE"]


def test_packed_store_on_emulator():
    """f4: 2-bit packed store (layout of main.cu:78-86 + validity bitmap): pack, unpack, count in chunks"""
    run_case("packed", 5, 100_003, 1)
    run_case("packed", 12, 250_017, 2, KC_PACKED_CHUNK=30_000)
    run_case("packed", 3, 1000, 3, KC_PACKED_CHUNK=64)
    for n in (31, 3, 16, 17):
        run_case("packed", 2, n, 6)


def test_host_packed_count_on_emulator():
    """kc_count_dense_host_packed: packer threads -> pinned ring -> H2D -> unpack -> count behind the copies.
    64-base items: a 200 K input goes through ~200 slots (25 ring wrap-arounds) and 13 count calls"""
    run_case("dense_host_packed", 5, 200_003, 1, "dirty", 3, KC_HOSTPACK_ITEM=64)
    run_case("dense_host_packed", 12, 300_000, 2, "genome", 2, KC_HOSTPACK_ITEM=96)
    run_case("dense_host_packed", 8, 70_001, 3, "genome", 1, KC_HOSTPACK_ITEM=32)
    run_case("dense_host_packed", 4, 5_000_000, 4, "dirty", 0)            # production item size, auto threads
    run_case("dense_host_packed", 6, 100_000, 6, "polyA", 2, KC_HOSTPACK_ITEM=64)   # no invalid byte: no mask words sent
    # sparse bitmap transfer: a few dirty blocks per slot (compaction on the host, rank lookup in mask_expand_kernel),
    # at 1-word blocks (item 64), 2-word blocks (item 2048) and mixed sparse / full slots
    run_case("dense_host_packed", 7, 150_000, 7, "sparseN300", 3, KC_HOSTPACK_ITEM=64)
    run_case("dense_host_packed", 7, 150_000, 8, "sparseN5000", 2, KC_HOSTPACK_ITEM=64)
    run_case("dense_host_packed", 9, 600_000, 9, "sparseN2000", 3, KC_HOSTPACK_ITEM=2048)
    run_case("dense_host_packed", 5, 200_000, 10, "sparseN100", 2, KC_HOSTPACK_ITEM=96)
    for n in (0, 1, 3, 4, 31, 32, 33):
        run_case("dense_host_packed", 3, n, 5, "dirty", 2, KC_HOSTPACK_ITEM=32)


@pytest.mark.parametrize("case", [(12, 300_000, "genome", 3, 0), (21, 200_000, "dirty", 4, 5), (31, 150_000, "sparseN40", 5, 9),
                                  (5, 100_000, "dirty", 6, 1)], ids=lambda c: "k%d-%s" % (c[0], c[2]))
def test_fingerprint_self_checks_on_emulator(case):
    """csrc/check.cu: window fingerprint of the input == fingerprint of its counts == the oracle's value"""
    run_case("fingerprint", *case)


def test_gpu_fasta_parser_on_emulator():
    """f2, device side: kc_import_seqs_device == the host loader on the reference-generated fixtures and
    random files, with 16- and 5-byte tiles (lines, ids and records span tiles; > 1024 tiles in one file)"""
    run_case("ingest", 1, 60, KC_INGEST_TILE=16)
    run_case("ingest", 2, 40, KC_INGEST_TILE=5)
    run_case("ingest", 3, 20)


@pytest.mark.skipif(not os.environ.get("KC_RUN_SLOW"), reason="45 s; set KC_RUN_SLOW=1 (run by hand after editing tests/test_gpu_parity_variants.py)")
def test_variant_gpu_cases_on_emulator():
    """the GPU-only cases of tests/test_gpu_parity_variants.py themselves, inputs shrunk 500x, on the emulator library with
    torch's CUDA surface replaced by CPU stand-ins: their Python has then run before it meets a B200"""
    env = dict(os.environ, KC_FIRST_RUN_SCALE="0.002", KC_SPARSE_RADIX_SHAPE="small", KC_EMU_SMS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_under_shim.py"), "-m", "pytest",
                        os.path.join(ROOT, "tests", "test_gpu_parity_variants.py"), "-m", "gpu", "-q", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "18 passed, 2 skipped" in r.stdout
