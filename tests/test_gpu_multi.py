"""Parity of the multi-GPU join (include/dwj.h: dwj_xj_*, dwj_mg_*; csrc/dwj_xj.cu) with the CPU oracle.

`virtual`: the single-process driver dwj_mg_* with several ranks that all live on cuda:0 -- the complete C++
orchestration (partition passes into send slots, counts and flags through "peer" memory, one host thread per rank,
kernels pulling through segment lists) runs on the one-GPU box the driver's GPU tests use.
`ranks`: one process per GPU under torchrun over torch symmetric memory (skipped when the box has a single GPU).
Both compare the concatenated result rows of several consecutive joins with oracle.sort_join, bit for bit.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mg_worker.py")


def _env(regions: bool):
    env = dict(os.environ)
    # several ranks on one GPU wait for each other inside kernels: every stream needs its own hardware queue
    env["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"
    env["DWJ_XJ_TIMEOUT_MS"] = "8000"              # a lost flag fails the test in seconds
    env["DWJ_TEST_WATCHDOG"] = "90"
    if regions:                      # force the table-region machinery at test size
        env["DWJ_PARTITION_MIN_MB"] = "0"
        env["DWJ_REGION_MB"] = "0.0625"
    return env


@pytest.mark.parametrize("world,key_bytes,pull,passes,regions", [
    (1, 4, "direct", 1, True), (2, 4, "direct", 1, True), (4, 8, "direct", 1, True), (2, 8, "scatter", 1, True),
    (4, 4, "scatter", 2, True), (2, 4, "direct", 2, False), (8, 8, "direct", 1, False), (1, 8, "direct", 4, True)])
def test_mg_join_virtual_ranks(tmp_path, world, key_bytes, pull, passes, regions):
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    out = tmp_path / "res.json"
    r = subprocess.run([sys.executable, WORKER, "virtual", str(world), str(key_bytes), pull, str(passes), str(out)],
                       env=_env(regions), capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert len(res["steps"]) == 3 and all(s["same"] and s["rows"] == s["want"] for s in res["steps"]), res
    assert res["info"]["direct_pull"] == (1 if pull == "direct" else 0)
    if world <= 4:
        assert res["info"]["chunks"] > res["info"]["ring"]          # the send slots are reused inside a join
    if regions:
        assert res["info"]["regions"] > 1


@pytest.mark.parametrize("world,key_bytes,passes", [(2, 4, 1), (4, 8, 2)])
def test_mg_join_virtual_ranks_duplicate_build_keys(tmp_path, world, key_bytes, passes):
    """Engines without the unique-keys flag behind the exchange: every build key three times (one-to-many tables on the
    receivers, rows grouped by table region by the pulling scatter), three consecutive joins against the oracle."""
    out = tmp_path / "res.json"
    env = _env(True)
    env["DWJ_TEST_DUP"] = "3"
    r = subprocess.run([sys.executable, WORKER, "virtual", str(world), str(key_bytes), "scatter", str(passes), str(out)],
                       env=env, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert len(res["steps"]) == 3 and all(s["same"] and s["rows"] == s["want"] for s in res["steps"]), res
    assert res["info"]["direct_pull"] == 0


@pytest.mark.parametrize("key_bytes,passes", [(8, 2), (4, 4)])
def test_one_gpu_passes_compact_then_partition(tmp_path, key_bytes, passes):
    """world == 1, passes > 1, one probe chunk: the local path of dwj_xj_join (dwj_filter_rows compacts a relation's key
    class, the compact copy is partitioned by table region, the slots serve as each other's scratch) -- the way the
    2^31 x 2^31 join runs on one GPU.  Three consecutive joins against the oracle."""
    out = tmp_path / "res.json"
    env = _env(True)
    env["DWJ_TEST_CHUNK_ROWS"] = "1000000"
    r = subprocess.run([sys.executable, WORKER, "virtual", "1", str(key_bytes), "direct", str(passes), str(out)],
                       env=env, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert all(s["same"] and s["rows"] == s["want"] for s in res["steps"]), res
    assert res["info"]["chunks"] == 1 and res["info"]["passes"] == passes and res["info"]["regions"] > 1


@pytest.mark.parametrize("key_bytes,pull,passes", [(4, "direct", 1), (8, "direct", 1), (8, "scatter", 1), (4, "scatter", 2)])
def test_pull_exchange_two_ranks(tmp_path, oracle, key_bytes, pull, passes):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = _env(True)
    env.pop("CUDA_DEVICE_MAX_CONNECTIONS")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), WORKER, "ranks", str(key_bytes), pull, str(passes), str(tmp_path)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-6000:]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mg_worker
    info = json.load(open(tmp_path / "info.json"))
    assert info["direct_pull"] == (1 if pull == "direct" else 0) and info["regions"] > 1
    for step in range(mg_worker.STEPS):
        ak, av, bk, bv = mg_worker.make_inputs(step, key_bytes, 400_000, 1_000_003)
        want = oracle.sort_join(ak, av, bk, bv)
        parts = [np.load(tmp_path / f"rank{r_}_step{step}.npz") for r_ in range(2)]
        got = pyoracle.canonical_rows(*(np.concatenate([p[c] for p in parts]) for c in ("k", "b", "p")))
        assert len(got[0]) == len(want[0])
        for w, x in zip(want, got):
            np.testing.assert_array_equal(w, x)
