"""CPU, world_size 2, gloo: the multi-GPU host logic — contiguous image sharding and the single
broadcast of the shared table blob — with the data-path collective count being exactly one."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_and_are_disjoint():
    sys.path.insert(0, ROOT)
    from jpegsnoop_b200.shard import shard_range
    for n in (0, 1, 7, 512, 1024, 4096, 4097):
        for world in (1, 2, 4, 8):
            rs = [shard_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(4096, 8, 3) == (1536, 2048)          # BASELINE config 3: image i on GPU floor(i/512)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import jpeg_cases as JC
    from jpegsnoop_b200 import BatchDecoder
    from jpegsnoop_b200.shard import shard_range, broadcast_tables, tables_to_bytes
    cases = [j for _, j in JC.small_cases()[:6]]
    lo, hi = shard_range(len(cases), world, rank)
    mine = cases[lo:hi]
    # every rank parses only its own shard; rank 0's table sets are broadcast (here all images of the
    # "shared-table" batch are forced onto rank 0's first set to model configs 2/3/5)
    tarr0, darr0, bits0 = BatchDecoder.prepare([cases[0]] * 2)
    tarr = broadcast_tables(tarr0 if rank == 0 else None, src=0)
    assert tables_to_bytes(tarr) == tables_to_bytes(tarr0), "table blob differs after broadcast"
    tarr_m, darr_m, bits_m = BatchDecoder.prepare(mine)
    assert len(darr_m) == hi - lo and all(d.scan_length > 0 for d in darr_m)
    total = [None] * world
    dist.all_gather_object(total, (lo, hi, int(sum(d.dim_x * d.dim_y for d in darr_m))))
    if rank == 0:
        assert [t[:2] for t in total] == [shard_range(len(cases), world, r) for r in range(world)]
        print("OK", total)
    dist.destroy_process_group()
""") % (ROOT, ROOT)


def test_two_rank_gloo_sharding_and_table_broadcast(built, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "OK" in r.stdout
