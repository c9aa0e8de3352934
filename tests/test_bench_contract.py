"""CPU: host-side pieces of bench.py that need no GPU — the algorithmic-bytes formula of SURVEY.md §8d, the kernel-source
hash that ties profiles/traffic.json to the kernels being timed, and the file rendezvous the ranks of a multi-GPU run use
instead of torch.distributed (three real processes)."""
import json
import multiprocessing as mp
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_the_survey_table():
    import bench

    assert bench.b_alg(1, 4) == 944.0 and bench.B_ALG == 944.0        # single star, V J H K
    assert bench.b_alg(2, 4) == 1848.0                                 # binary, 4 bands
    assert bench.b_alg(1, 11) == 1840.0                                # single star, 11 default bands
    assert len(bench.BANDS11) == 11


def test_traffic_file_is_tied_to_kernel_sources():
    import bench

    h = bench.kernel_source_hash()
    assert len(h) == 16 and h == bench.kernel_source_hash()
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        pytest.skip("no ncu capture committed yet")
    with open(path) as f:
        tj = json.load(f)
    assert set(tj) >= {"kernel_source_hash", "batches"} and {"posterior_like", "grid_wide"} <= set(tj["batches"])
    for b in tj["batches"].values():
        assert b["dram_bytes_per_launch"] > 0 and b["algorithmic_bytes_per_launch"] == 944_000_000
        assert "iso_lnpost_kernel" in b["kernel"] and os.path.exists(os.path.join(ROOT, b["capture"]))
    # the HBM-bound batch must not re-read: its DRAM traffic stays below the algorithmic bytes
    assert tj["batches"]["grid_wide"]["dram_bytes_per_launch"] < 944_000_000


def _rank(path, rank, world, q):
    from isochrones_b200.parallel import FileRendezvous

    rz = FileRendezvous(path, rank, world, timeout=60.0)
    got = rz.allgather_bytes(b"payload-%d" % rank)
    first = rz.broadcast(b"from-zero" if rank == 0 else b"ignored")
    big = rz.max(10.0 + rank)
    every = rz.all(True), rz.all(rank != 1)
    rz.barrier()
    rz.close()
    q.put((rank, got, first, big, every))


def test_file_rendezvous_three_processes(tmp_path):
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(str(tmp_path / "rdzv"), r, world, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, first, big, every in res:
        assert got == [b"payload-0", b"payload-1", b"payload-2"] and first == b"from-zero"
        assert big == 12.0 and every == (True, False)
    assert not os.path.exists(str(tmp_path / "rdzv"))          # rank 0 removed the directory once all had signed off
