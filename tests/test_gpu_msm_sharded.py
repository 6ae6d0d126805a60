"""Multi-GPU MSM (cdl_g1_msm_sharded), windows dealt to the ranks or points partitioned:
one process per GPU, NCCL all-gather of one partial sum per rank.  Needs >= 2 GPUs (run with
`gpurun --gpus 2`); skipped on a single-GPU box, where
test_gpu_msm_big.test_window_partition_partials_add_up covers the arithmetic."""
import importlib
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, uid, n, q, partition=""):
    if partition:
        os.environ["CDL_MSM_PARTITION"] = partition  # "windows" / "points": force one split (default: by window count)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    pkg = importlib.import_module("go-curdleproofs_b200")
    from oracle import bls12381 as b
    from oracle.rand import Rand
    from util import R, aff_enc, frs_enc, jac_dec

    ctx = pkg.Context(rank)
    ctx.comm_init(uid, rank, world)
    a = Rand(5).get_frs(n)
    s = Rand(6).get_frs(n)
    pts = ctx.g1_scalar_mul_affine(aff_enc(b.G1_GEN) * n, frs_enc(a), broadcast=False)
    out = ctx.g1_msm_sharded(pts, frs_enc(s))
    want = b.g1_mul(b.G1_GEN, sum(x * y for x, y in zip(a, s)) % R)
    q.put((rank, jac_dec(out) == want, out))
    ctx.comm_destroy()
    ctx.close()


@pytest.mark.parametrize("partition", ["windows", "points"])
@pytest.mark.parametrize("n", [4096, 50000])
def test_sharded_msm_two_ranks(pkg, n, partition):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    uid = pkg.comm_unique_id()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, uid, n, q, partition)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert len({out for _, _, out in res}) == 1  # every rank holds the same normalised point
