"""Multi-GPU operator layer on whatever the box has.

Single-GPU box: `world` processes share cuda:0 and exchange over gloo (device buffers staged through the host) - slow, but
every exchange point of the C++ operators runs for real: partial-state all-gather + ordered merge, bitmap-join sum,
broadcast join, key-hash shuffle, consistent errors, empty shards.  With >= 2 GPUs the same suite also runs over NCCL."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    results = []
    try:
        sys.path.insert(0, ROOT)
        from tests import dist_sql
        results = dist_sql._Log(rank)
        dist_sql.run_rank(rank, world, "gloo", results)
    except Exception:  # noqa: BLE001
        import traceback
        results.append(("worker", traceback.format_exc()[-2000:]))
    finally:
        q.put((rank, list(results)))
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sql_matches_oracle_shared_gpu(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    for rank, results in out:
        bad = [r for r in results if r[1] != "ok"]
        assert results and not bad, f"rank {rank}: {bad[:3]}"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs for NCCL")
@pytest.mark.parametrize("exchange", ["native", "callbacks"])
def test_sharded_sql_matches_oracle_nccl(exchange):
    """native: the library's own NCCL collectives (bq_comm_*); callbacks: torch.distributed behind the exchange table."""
    n = min(4, torch.cuda.device_count())
    n = 1 << (n.bit_length() - 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_sql.py")],
                       capture_output=True, text=True, timeout=1500, cwd=ROOT, env=dict(os.environ, BOSQL_TEST_EXCHANGE=exchange))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
