"""Multi-rank plumbing on CPU (gloo, world_size 2): shard generation, the exchange-vector protocol (one sum
all-reduce that doubles as an all-gather of per-rank lanes) and the cluster-wide scalars every rank derives
from it, checked against the oracle on the unsharded cluster. The CUDA side of the same protocol is
ust_kernels.cu:derive_scalars; the NCCL path itself is covered by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from helpers import abi
from ust import synth

V_HIST, V_UNAV, V_CAND, MAXW = 0, 16, 17, 8
V_RANK_CAND, V_RANK_NODES, V_RANK_ERRINV = 18, 18 + MAXW, 18 + 2 * MAXW
V_LEN = 18 + 3 * MAXW
PASS_OF_STATE = [0, 2, 3, 4, 5, 6, 7, -1, 8, 10, 11, 1, 9, -1, -1, -1]


def local_vector(soa, rank):
    """What one shard contributes (mirror of load_local_vector / the streaming phase's counters)."""
    code = soa["state"] & 15
    v = np.zeros(V_LEN, np.int64)
    for c in range(16):
        v[V_HIST + min(c, 14)] += int(np.sum(code == c))
    inside = code < 14
    v[V_UNAV] = int(np.sum(inside & ((soa["state"] & (abi.UST_HOT_UNSCHEDULABLE | abi.UST_HOT_NOT_READY)) != 0)))
    cand = int(np.sum((code == 1) & ((soa["state"] & abi.UST_HOT_SKIP) == 0)))
    v[V_CAND] = cand
    v[V_RANK_CAND + rank] = cand
    v[V_RANK_NODES + rank] = soa["state"].shape[0]
    err = ((soa["state"] & abi.UST_HOT_REVISION_HASH_ERROR) != 0) & ((soa["flags"] & abi.UST_F_POD_ORPHANED) == 0) & \
        np.isin(code, [0, 8, 11, 12])
    if err.any():
        idx = np.nonzero(err)[0]
        keys = [(PASS_OF_STATE[int(code[i])] << 56) | (int(i) + 1) for i in idx]
        v[V_RANK_ERRINV + rank] = np.array(~np.uint64(min(keys))).view(np.int64)
    return v


def derive(v, world):
    """Mirror of derive_scalars for the lanes every rank needs."""
    offs = np.concatenate([[0], np.cumsum(v[V_RANK_NODES:V_RANK_NODES + world])])
    cand_before = np.concatenate([[0], np.cumsum(v[V_RANK_CAND:V_RANK_CAND + world])])
    abort = None
    for r in range(world):
        e = np.uint64(np.array(v[V_RANK_ERRINV + r]).view(np.uint64))
        if e:
            k = int(~e & np.uint64(0xFFFFFFFFFFFFFFFF))
            gk = (k & (0xFF << 56)) | ((k & ((1 << 56) - 1)) + int(offs[r]))
            abort = gk if abort is None else min(abort, gk)
    return offs, cand_before, abort


def _worker(rank, world, port, n, seed, p_err, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    soa = synth.make_nodes(n, seed, start=rank * n, error_pct=p_err)
    v = torch.from_numpy(local_vector(soa, rank))
    dist.all_reduce(v, op=dist.ReduceOp.SUM)  # one collective per ApplyState
    offs, cand_before, abort = derive(v.numpy(), world)
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: soa[k] for k in ("state", "flags", "pod_rev", "ds_idx")})
    if rank == 0:
        q.put((v.numpy().copy(), offs, cand_before, abort, gathered))
    dist.destroy_process_group()


@pytest.mark.parametrize("p_err", [0.0, 0.01])
def test_exchange_vector_protocol_world2(p_err):
    world, n, seed = 2, 20_000, 0x5EED0005
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, p_err, q)) for r in range(world)]
    for p in procs:
        p.start()
    v, offs, cand_before, abort, shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards are contiguous pieces of ONE cluster: same bytes as generating it whole
    whole = synth.make_nodes(world * n, seed, error_pct=p_err)
    for k in ("state", "flags", "pod_rev", "ds_idx"):
        assert np.array_equal(np.concatenate([s[k] for s in shards]), whole[k]), k
    pol = synth.config_policy("C3")
    rc, nxt, act, oc, cnt = helpers.oracle_apply(pol, whole, variant=1)
    assert list(v[V_HIST:V_HIST + 16]) == cnt["hist"]
    assert v[V_UNAV] == cnt["unavailable"] and v[V_CAND] == cnt["candidates"]
    assert list(offs) == [0, n, 2 * n]
    assert cand_before[2] == cnt["candidates"]
    if rc:
        assert abort is not None and (abort >> 56) == cnt["error_pass"] and (abort & ((1 << 56) - 1)) - 1 == cnt["error_index"]
    else:
        assert abort is None


def test_pod_lists_shard_at_node_boundaries():
    """Sharded generation (what each rank of tests/test_gpu_multi.py and bench.py feeds its GPU) is the unsharded
    snapshot cut at node boundaries: per-rank CSR offsets start at 0 and the concatenated lists are the whole."""
    import numpy as np
    from ust import synth
    n, world, seed = 3000, 3, 0x5EED0004
    whole_nodes = synth.make_nodes(world * n, seed)
    whole_pods = synth.make_pods(world * n, seed)
    flags, total = [], 0
    for r in range(world):
        part = synth.make_nodes(n, seed, start=r * n)
        for k in ("state", "flags", "pod_rev", "ds_idx"):
            assert np.array_equal(part[k], whole_nodes[k][r * n:(r + 1) * n]), k
        pods = synth.make_pods(n, seed, start=r * n)
        assert pods["pod_off"][0] == 0
        lo, hi = int(whole_pods["pod_off"][r * n]), int(whole_pods["pod_off"][(r + 1) * n])
        assert np.array_equal(pods["pod_off"], whole_pods["pod_off"][r * n:(r + 1) * n + 1] - lo)
        assert np.array_equal(pods["pod_flags"], whole_pods["pod_flags"][lo:hi])
        total += hi - lo
    assert total == whole_pods["pod_flags"].shape[0]
    blocked = synth.make_pods_blocked(world * n, seed, block=777)
    assert np.array_equal(blocked["pod_off"], whole_pods["pod_off"]) and np.array_equal(blocked["pod_flags"], whole_pods["pod_flags"])
