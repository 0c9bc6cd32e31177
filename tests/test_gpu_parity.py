"""Parity of the CUDA path (through the C ABI of libust.so) against the oracle. Bit-exact: this is integer
and index work. Run on the B200 box: python -m pytest tests -m gpu"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from helpers import abi
from ust import lib as ustlib, synth

pytestmark = pytest.mark.gpu

G = helpers.load_golden()


@pytest.fixture(scope="module")
def handle():
    h = ustlib.Handle(0)  # raises (never skips) when the extension or the device is missing
    yield h
    h.close()


def gpu_apply(handle, pol, soa, pods=None, nil_policy=False):
    return handle.apply_state(None if nil_policy else pol, soa, pods, want_outcome=True)


# ---- the reference's own known-answer tests, through the kernel ------------------------------------

VEC = [v for v in G["apply_state"] if not v.get("nil_state")]


@pytest.mark.parametrize("v", VEC, ids=[v["ref"].split("/")[-1] for v in VEC])
def test_reference_vector(handle, v):
    pol = helpers.policy_from_vector(v)
    soa, pods = helpers.encode_nodes(v["nodes"], G["daemonset_hash"], v.get("policy"))
    rc, nxt, act, oc, cnt = gpu_apply(handle, pol, soa, pods, nil_policy=pol is None)
    assert helpers.check_vector(v, soa, rc, nxt, act, oc) > 0 or v.get("actuator_error")
    helpers.check_derived(v, nxt)
    # and identical to the oracle in every output, not only in what the Go test asserts
    ref = helpers.oracle_apply(pol if pol is not None else abi.Policy(), soa, pods, variant=0, nil_policy=pol is None)
    helpers.assert_same((rc, nxt, act, oc, cnt), ref, v["name"])


def test_nil_state_is_an_error(handle):
    # upgrade_state.go:175-177
    pol = abi.make_policy()
    rc = ustlib.load().ust_apply_state(handle._h, C.addressof(pol), 4, None, None, None, None, 0, None, None, None, None, None, None)
    assert rc == abi.UST_ERR_NIL_STATE
    assert "currentState should not be empty" in handle.last_error()


@pytest.mark.parametrize("b", G["build_state"], ids=lambda b: b["name"][:40])
def test_build_state_vector(handle, b):
    pods = b["pods"]
    n = len(pods)
    state = np.zeros(max(n, 1), np.uint8)[:n]
    ds_idx = np.full(max(n, 1), -1, np.int32)[:n]
    for i, p in enumerate(pods):
        code = abi.STATE_CODE.get(p["node_state"], abi.UST_STATE_OTHER)
        if p["node_name"] == "" and p["phase"] == "Pending":
            code = abi.UST_STATE_EXCLUDED
        state[i] = code
        ds_idx[i] = -1 if p["ds"] is None else p["ds"]
    desired = np.array([d["desired"] for d in b["daemonsets"]], np.int32)
    rc, cnt = handle.build_state(state, ds_idx, desired)
    ocnt = abi.Counters()
    orc = helpers.oracle().ust_oracle_build_state(
        C.c_int64(n), state.ctypes.data_as(C.c_void_p), ds_idx.ctypes.data_as(C.c_void_p), C.c_int32(len(desired)),
        desired.ctypes.data_as(C.c_void_p), C.byref(ocnt))
    assert rc == orc
    if b["expect_error"]:
        assert rc == abi.K["UST_ERR_" + b["expect_error"]]
    else:
        assert rc == 0
        got = {abi.STATE_NAMES[c]: cnt["hist"][c] for c in range(13) if cnt["hist"][c]}
        assert got == b["expect_buckets"]
    assert cnt == ocnt.as_dict()


@pytest.mark.parametrize("b", G["build_state"], ids=lambda b: b["name"][:40])
def test_build_state_vector_uid_join(handle, b):
    """The same specs with the owner join done on the device from 128-bit owner UIDs."""
    state, owner, ds_uid, desired = helpers.uid_inputs_from_vector(b, np.random.default_rng(11))
    rc, ds_idx, cnt = handle.build_state_uids(state, owner, ds_uid, desired)
    orc, ods, ocnt = helpers.oracle_build_state_uids(state, owner, ds_uid, desired)
    assert rc == orc and cnt == ocnt and np.array_equal(ds_idx, ods)
    if b["expect_error"]:
        assert rc == abi.K["UST_ERR_" + b["expect_error"]]
    else:
        assert rc == 0
        assert {abi.STATE_NAMES[c]: cnt["hist"][c] for c in range(13) if cnt["hist"][c]} == b["expect_buckets"]


@pytest.mark.parametrize("n,n_ds", [(1, 1), (257, 2), (70_001, 7), (1_000_003, 4), (300_000, 1024), (200_000, 3000)])
def test_build_state_uid_join_random(handle, n, n_ds):
    """Owned / orphaned / foreign-owned driver pods in random order, few and many DaemonSets (shared-memory table and
    the global-memory fallback), with and without a DaemonSet that misses pods (upgrade_state.go:128-131)."""
    rng = np.random.default_rng(n + n_ds)
    ds_uid = rng.integers(1, 2 ** 63, size=(n_ds, 2), dtype=np.uint64)
    kind = rng.random(n)
    truth = np.where(kind < 0.05, -1, np.where(kind < 0.12, -2, rng.integers(0, n_ds, n))).astype(np.int32)
    owner = np.zeros((n, 2), np.uint64)
    owner[truth >= 0] = ds_uid[truth[truth >= 0]]
    foreign = truth == -2
    owner[foreign] = rng.integers(2 ** 63, 2 ** 64 - 1, size=(int(foreign.sum()), 2), dtype=np.uint64)
    # near misses: same first half as a real DaemonSet, different second half
    nm = np.flatnonzero(foreign)[::3]
    owner[nm, 0] = ds_uid[rng.integers(0, n_ds, nm.shape[0]), 0]
    state = (rng.integers(0, 16, n).astype(np.uint8) | (rng.integers(0, 16, n).astype(np.uint8) << 4)).astype(np.uint8)
    for miss in (False, True):
        desired = np.bincount(truth[truth >= 0], minlength=n_ds).astype(np.int32)
        if miss:
            desired[n_ds // 2] += 1
        rc, ds_idx, cnt = handle.build_state_uids(state, owner, ds_uid, desired)
        orc, ods, ocnt = helpers.oracle_build_state_uids(state, owner, ds_uid, desired)
        assert rc == orc == (abi.K["UST_ERR_DS_UNSCHEDULED"] if miss else 0)
        assert np.array_equal(ds_idx, ods) and np.array_equal(ds_idx, truth)
        assert cnt == ocnt
    dup = np.concatenate([ds_uid[:1], ds_uid[:1]])
    assert handle.build_state_uids(state[:1], owner[:1], dup, np.zeros(2, np.int32))[0] == abi.K["UST_ERR_INVALID_ARGUMENT"]


# ---- random snapshots: every flag bit, abort paths, requestor mode, pod lists ------------------------

SIZES = [0, 1, 3, 4, 5, 127, 128, 129, 1023, 1024, 1025, 4095, 4096, 4097, 8191, 8192, 20_000, 70_001, 300_000]


@pytest.mark.parametrize("n", SIZES)
def test_random_snapshot_matches_reference_shaped_oracle(handle, n):
    rng = np.random.default_rng(4242 + n)
    for rep in range(3):
        soa, pods = helpers.random_soa(rng, n, n_ds=int(rng.integers(1, 6)), p_err=(0.0, 0.0005, 0.02)[rep],
                                       with_pods=(rep == 1))
        pol = helpers.random_policy(rng)
        got = gpu_apply(handle, pol, soa, pods)
        ref = helpers.oracle_apply(pol, soa, pods, variant=0)
        helpers.assert_same(got, ref, f"n={n} rep={rep}")


@pytest.mark.parametrize("seed", range(12))
def test_slot_budget_cut_positions(handle, seed):
    """The ordered slot allocation (upgrade_inplace.go:71-109) with the budget cut falling at the start,
    inside and at the end of a chunk, with maxUnavailable percent / int / nil."""
    rng = np.random.default_rng(seed)
    n = int(rng.choice([9000, 33_000, 150_000, 600_000]))
    soa, _ = helpers.random_soa(rng, n, all_states=False, wild=False)
    pick = rng.random(n)
    soa["state"] = np.where(pick < 0.6, (soa["state"] & 0xF0) | 1, soa["state"]).astype(np.uint8)
    soa["state"] &= np.uint8(0x7F)
    cands = int(np.sum(((soa["state"] & 15) == 1) & ((soa["state"] & abi.UST_HOT_SKIP) == 0)))
    for max_par, unav in ((0, None), (0, "100%"), (0, f"{int(rng.integers(20, 60))}%"), (int(rng.integers(1, 200)), None),
                          (cands // 2 + n, int(n)), (0, int(rng.integers(0, n))), (1, 0), (10**9, "37%")):
        pol = abi.make_policy(max_parallel_upgrades=max_par, max_unavailable=unav)
        got = gpu_apply(handle, pol, soa)
        ref = helpers.oracle_apply(pol, soa, variant=1)
        helpers.assert_same(got, ref, f"seed={seed} maxPar={max_par} maxUnav={unav}")


@pytest.mark.parametrize("n", [300_000, 1_000_037, 2_500_063])
def test_speculation_hint_never_changes_results(handle, n):
    """The kernel speculates where the slot budget cuts and carries the observed cut to the next call with the
    same size and policy. Whatever the hint (none, exact, stale because the snapshot changed, cut moved to the
    ragged end), the grants must be the ordered ones of upgrade_inplace.go:71-109."""
    pol = abi.make_policy(max_parallel_upgrades=0, max_unavailable="30%")
    for rep, frac in enumerate((0.6, 0.6, 0.95, 0.31, 0.02, 0.6)):
        rng = np.random.default_rng(4000 + rep if rep != 1 else 4000)   # rep 1 repeats rep 0: exact hint
        soa, _ = helpers.random_soa(rng, n, all_states=False, wild=False)
        pick = rng.random(n)
        soa["state"] = np.where(pick < frac, (soa["state"] & 0xF0) | 1, soa["state"]).astype(np.uint8)
        soa["state"] &= np.uint8(0x7F)
        got = gpu_apply(handle, pol, soa)
        ref = helpers.oracle_apply(pol, soa, variant=1)
        helpers.assert_same(got, ref, f"n={n} rep={rep} frac={frac}")


@pytest.mark.parametrize("seed", range(6))
def test_pipelined_host_path(handle, seed):
    """Snapshots >= 2^19 nodes take the segmented upload/compute/download pipeline of ust_apply_state; the
    speculative outputs of early segments must be replaced when the verification redoes chunks (aborts, budgets
    that cut through the middle, requestor mode)."""
    rng = np.random.default_rng(700 + seed)
    n = 700_001 + 1000 * seed
    soa, _ = helpers.random_soa(rng, n, p_err=(0.0, 1e-6, 1e-5)[seed % 3], wild=bool(seed % 2))
    pol = helpers.random_policy(rng)
    if seed == 3:
        pol = abi.make_policy(max_parallel_upgrades=0, max_unavailable="30%")
    if seed == 4:
        pol = abi.make_policy(max_parallel_upgrades=60_000)
    got = gpu_apply(handle, pol, soa)
    ref = helpers.oracle_apply(pol, soa, variant=1)
    helpers.assert_same(got, ref, f"pipelined seed={seed}")


@pytest.mark.parametrize("seed", range(3))
def test_long_and_empty_pod_lists(handle, seed):
    """Pod lists of every shape the list walker distinguishes: empty, shorter than one 16-byte chunk, unaligned
    starts, more than one pass (> 48 pods), and the last list ending exactly at the end of the array."""
    rng = np.random.default_rng(900 + seed)
    n = (4096 * 3 + 17, 60_001, 4096)[seed]
    soa, _ = helpers.random_soa(rng, n, with_pods=False, wild=bool(seed))
    shape = rng.random(n)
    cnt = np.where(shape < 0.3, 0, np.where(shape < 0.7, rng.integers(1, 9, n), np.where(shape < 0.95, rng.integers(9, 49, n),
                                                                                         rng.integers(49, 300, n))))
    if seed == 2:
        cnt[-1] = 0   # ... and an empty list at the very end
    off = np.zeros(n + 1, np.int32)
    np.cumsum(cnt, out=off[1:])
    total = int(off[-1])
    pf = rng.integers(0, 6, size=total).astype(np.uint16)
    for k, v in abi.K.items():
        if k.startswith("UST_POD_") and k != "UST_POD_PHASE_MASK":
            pf |= np.where(rng.random(total) < 0.08, np.uint16(v), np.uint16(0))   # sparse: long lists are not all-ones
    pods = {"pod_off": off, "pod_flags": pf}
    for rep in range(3):
        pol = helpers.random_policy(rng)
        pol.evaluate_actuators = 1
        got = gpu_apply(handle, pol, soa, pods)
        ref = helpers.oracle_apply(pol, soa, pods, variant=0)
        helpers.assert_same(got, ref, f"seed={seed} rep={rep}")


@pytest.mark.parametrize("n", [5000, 700_001])
def test_delta_updates_of_the_resident_snapshot(handle, n):
    """ust_apply_state_delta: re-encode a few nodes, evaluate the resident snapshot again - identical to a full call
    on the updated arrays (direct and pipelined upload paths leave the same resident arrays), any policy, repeated."""
    rng = np.random.default_rng(n)
    soa, _ = helpers.random_soa(rng, n, wild=True)
    pol = helpers.random_policy(rng)
    helpers.assert_same(gpu_apply(handle, pol, soa), helpers.oracle_apply(pol, soa, variant=1), "full call")
    for rep, frac in enumerate((0.0, 0.0005, 0.02, 0.3)):
        m = int(n * frac)
        idx = rng.choice(n, size=m, replace=False).astype(np.int64)
        fresh, _ = helpers.random_soa(rng, m, wild=True, p_err=1e-4 if rep == 3 else 0.0)
        for k in ("state", "flags", "pod_rev", "ds_idx"):
            soa[k][idx] = fresh[k]
        if rep == 2:
            soa["ds_rev"] = (soa["ds_rev"] + 1).astype(np.int32)   # a DaemonSet rolled to a new revision
        pol = helpers.random_policy(rng)
        got = handle.apply_state_delta(pol, n, idx, {k: fresh[k] for k in ("state", "flags", "pod_rev", "ds_idx")}, soa["ds_rev"])
        ref = helpers.oracle_apply(pol, soa, variant=1)
        helpers.assert_same(got, ref, f"delta rep={rep} m={m}")
    # contract errors
    bad = handle.apply_state_delta(pol, n, np.array([n], np.int64), {k: soa[k][:1] for k in ("state", "flags", "pod_rev", "ds_idx")}, soa["ds_rev"])
    assert bad[0] == abi.K["UST_ERR_INVALID_ARGUMENT"]
    handle.build_state(soa["state"][:10], np.zeros(10, np.int32), np.array([10], np.int32))   # shares the staging arrays
    gone = handle.apply_state_delta(pol, n, np.zeros(0, np.int64), {k: soa[k][:0] for k in ("state", "flags", "pod_rev", "ds_idx")}, soa["ds_rev"])
    assert gone[0] == abi.K["UST_ERR_INVALID_ARGUMENT"]


@pytest.mark.parametrize("n", [4097, 700_001])
def test_sparse_delta_outputs(handle, n):
    """ust_apply_state_delta_sparse: patching the previous call's outputs with the returned (index, next_state, actions)
    entries gives exactly the oracle's outputs on the updated snapshot - over several reconciles, with policy changes,
    a slot budget that moves, an abort, an overflow of the caller's arrays (then ust_fetch_outputs)."""
    rng = np.random.default_rng(31 + n)
    soa, _ = helpers.random_soa(rng, n, wild=True)
    pol = abi.make_policy(max_parallel_upgrades=0, max_unavailable="30%")
    rc, nxt, act, _, _ = handle.apply_state(pol, soa, want_outcome=False)
    helpers.assert_same((rc, nxt, act, None, None), helpers.oracle_apply(pol, soa, variant=1)[:3] + (None, None), "full call")
    cap = n // 4 + 16
    for rep, frac in enumerate((0.0, 0.001, 0.01, 0.05, 0.01, 0.6)):
        m = int(n * frac)
        idx = np.sort(rng.choice(n, size=m, replace=False)).astype(np.int64)
        fresh, _ = helpers.random_soa(rng, m, wild=True, p_err=2e-4 if rep == 4 else 0.0)
        for k in ("state", "flags", "pod_rev", "ds_idx"):
            soa[k][idx] = fresh[k]
        if rep == 3:
            pol = abi.make_policy(max_parallel_upgrades=int(n // 7), max_unavailable="55%")
        rc, n_out, oi, on, oa, cnt = handle.apply_state_delta_sparse(pol, idx, {k: fresh[k] for k in ("state", "flags", "pod_rev", "ds_idx")},
                                                                    soa["ds_rev"], cap)
        ref = helpers.oracle_apply(pol, soa, variant=1)
        expect_changed = int(np.sum((ref[1] != nxt) | (ref[2] != act)))
        assert n_out == expect_changed, (rep, n_out, expect_changed)
        if n_out > cap:
            assert rc == abi.K["UST_ERR_TRUNCATED"]
            frc, nxt, act = handle.fetch_outputs(n)
            assert frc == 0
        else:
            assert rc == ref[0], (rep, rc, ref[0])
            assert np.all(np.diff(oi[:n_out]) > 0)          # node order
            nxt[oi[:n_out]] = on[:n_out]
            act[oi[:n_out]] = oa[:n_out]
        assert np.array_equal(nxt, ref[1]) and np.array_equal(act, ref[2]), f"rep={rep}"
        assert cnt == ref[4] or n_out > cap
    handle.build_state(soa["state"][:10], np.zeros(10, np.int32), np.array([10], np.int32))   # drops the resident snapshot
    rc = handle.apply_state_delta_sparse(pol, np.zeros(0, np.int64), {k: soa[k][:0] for k in ("state", "flags", "pod_rev", "ds_idx")},
                                         soa["ds_rev"], 16)[0]
    assert rc == abi.K["UST_ERR_INVALID_ARGUMENT"]


@pytest.mark.parametrize("n,steps", [(3000, 25), (120_000, 30), (700_001, 8)])
def test_simulated_rollout_matches_the_cpu_simulation(handle, n, steps):
    """ust_simulate_rollout (ApplyState + feedback kernel, `steps` times on the device) against the oracle's
    simulation: every reconcile's counters and the final snapshot, bit for bit; then the error paths."""
    rng = np.random.default_rng(n + steps)
    soa = synth.make_nodes(n, 0xABC + n)
    soa["flags"] = (soa["flags"] | np.where(rng.random(n) < 0.2, np.uint32(abi.UST_F_WAIT_PODS_RUNNING), np.uint32(0))).astype(np.uint32)
    for pol in (abi.make_policy(max_parallel_upgrades=0, max_unavailable="10%"),
                abi.make_policy(max_parallel_upgrades=max(1, n // 50), max_unavailable=None, pod_deletion_enabled=True,
                                validation_enabled=True, pod_deletion={"force": True}, drain={"enable": True},
                                wait_for_completion={"podSelector": "app=job", "timeoutSeconds": 60})):
        assert gpu_apply(handle, pol, soa)[0] == 0          # makes the snapshot resident
        rc, done, hist, fin = handle.simulate_rollout(pol, n, steps)
        orc, odone, ohist, ofin = helpers.oracle_simulate(pol, soa, steps, variant=1)
        assert rc == orc == 0 and done == odone == steps
        for k in range(steps):
            assert hist[k] == ohist[k], f"counters of reconcile {k}"
        for key in ("state", "flags", "pod_rev"):
            assert np.array_equal(fin[key], ofin[key]), key
        # the snapshot stays resident: a delta call with nothing changed evaluates the simulated state
        got = handle.apply_state_delta(pol, n, np.zeros(0, np.int64), {k: soa[k][:0] for k in ("state", "flags", "pod_rev", "ds_idx")}, soa["ds_rev"])
        after = dict(soa); after.update(ofin)
        helpers.assert_same(got, helpers.oracle_apply(pol, after, variant=1), "ApplyState on the simulated snapshot")
    # a reconcile that fails (maxUnavailable does not parse) feeds nothing back
    bad = abi.make_policy(max_parallel_upgrades=1, max_unavailable="a-few")
    gpu_apply(handle, bad, soa)
    rc, done, hist, fin = handle.simulate_rollout(bad, n, 3)
    orc, odone, ohist, ofin = helpers.oracle_simulate(bad, soa, 3, variant=1)
    assert rc == orc == abi.K["UST_ERR_MAX_UNAVAILABLE"] and done == odone == 0
    assert all(np.array_equal(fin[k], soa[k]) for k in ("state", "flags", "pod_rev"))


@pytest.mark.parametrize("n,steps", [(4000, 40), (300_001, 16)])
def test_timed_and_requestor_rollout_simulation(handle, n, steps):
    """ust_simulate_rollout_timed (a clock: wait-for-completion and validation timeouts, jobs and validation pods that
    take time) and requestor mode, against the oracle's independent restatement: every reconcile's counters and the
    final snapshot, bit for bit."""
    rng = np.random.default_rng(n + steps)
    soa = synth.make_nodes(n, 0xDEF + n, requestor_pct=3.0)
    soa["flags"] = (soa["flags"] | np.where(rng.random(n) < 0.3, np.uint32(abi.UST_F_WAIT_PODS_RUNNING), np.uint32(0))
                    | np.where(rng.random(n) < 0.05, np.uint32(abi.UST_F_WAIT_START_ANNO), np.uint32(0))).astype(np.uint32)
    soa["flags"] = (soa["flags"] | np.where((rng.random(n) < 0.3) & ((soa["flags"] & abi.UST_F_WAIT_START_ANNO) != 0),
                                            np.uint32(abi.UST_F_WAIT_TIMED_OUT), np.uint32(0))).astype(np.uint32)
    wfc = {"podSelector": "app=job", "timeoutSeconds": 90}
    cases = [
        (abi.make_policy(max_parallel_upgrades=max(1, n // 40), max_unavailable="40%", validation_enabled=True, drain={"enable": True},
                         wait_for_completion=wfc), abi.SimOptions(30, 90, 200, 75, 600, 0)),
        (abi.make_policy(max_parallel_upgrades=0, max_unavailable="20%", validation_enabled=True, wait_for_completion=wfc),
         abi.SimOptions(60, 90, 45, -1, 300, 0)),                 # validation pods never become ready: nodes time out into upgrade-failed
        (abi.make_policy(max_parallel_upgrades=max(1, n // 40), validation_enabled=True, use_maintenance_operator=True,
                         wait_for_completion={"podSelector": "app=job"}), abi.SimOptions(20, 0, 50, 30, 600, 70)),
        (abi.make_policy(max_parallel_upgrades=5, use_maintenance_operator=True), None),   # requestor mode without a clock
    ]
    for pol, opt in cases:
        assert gpu_apply(handle, pol, soa)[0] == 0          # makes the snapshot resident
        if opt is None:
            rc, done, hist, fin = handle.simulate_rollout(pol, n, steps)
        else:
            rc, done, hist, fin = handle.simulate_rollout_timed(pol, opt, n, steps)
        orc, odone, ohist, ofin = helpers.oracle_simulate_timed(pol, opt, soa, steps, variant=1)
        assert rc == orc == 0 and done == odone == steps, (rc, orc, handle.last_error())
        for k in range(steps):
            assert hist[k] == ohist[k], f"counters of reconcile {k}"
        for key in ("state", "flags", "pod_rev"):
            assert np.array_equal(fin[key], ofin[key]), key
    # the clock does something: validation pods that never come up hold their nodes in validation-required until the
    # validation timeout (start annotation at t = 0, now > 0 + 300 first at t = 360 = reconcile 6), then the nodes fail
    v = synth.make_nodes(64, 0x77)
    v["state"] = ((v["state"] & 0xF0) | abi.UST_STATE_VALIDATION_REQUIRED).astype(np.uint8)
    v["flags"] = (v["flags"] & ~np.uint32(abi.UST_F_VALIDATION_DONE | abi.UST_F_SAFE_LOAD)).astype(np.uint32)
    pol, opt = cases[1]
    gpu_apply(handle, pol, v)
    rc, _, hist, _ = handle.simulate_rollout_timed(pol, opt, 64, 9)
    assert rc == 0
    assert [h["hist"][abi.UST_STATE_VALIDATION_REQUIRED] for h in hist] == [64] * 7 + [0, 0]
    bad = abi.SimOptions(30, 0, 0, 0, 600, 0)
    assert handle.simulate_rollout_timed(cases[0][0], bad, n, 1)[0] == abi.K["UST_ERR_INVALID_ARGUMENT"]   # timeout flag / value disagree


def test_simulated_wait_timeout_follows_the_reference_vectors(handle):
    """The wait-for-completion timeout of the simulation, pinned: a node whose wait-selector pods keep running is driven
    through the device simulation, and reconcile by reconcile the result must be what the reference's
    HandleTimeoutOnPodCompletions does (pod_manager.go:331-368; golden vectors pod_manager_test.go:183-229): no start
    annotation -> the annotation is set; within the timeout -> nothing; now > start + timeout -> pod-deletion-required
    and the annotation is gone. The expectation is computed here from the annotation rules alone, with the oracle's
    (vector-pinned) ApplyState evaluating each reconcile."""
    timeout, dt, steps = 100, 30, 8
    pol, soa, state, start = helpers.wait_timeout_timeline(G["daemonset_hash"], timeout, dt, steps)
    assert gpu_apply(handle, pol, soa)[0] == 0
    opt = abi.SimOptions(dt, timeout, 10 ** 6, 0, 600, 0)
    rc, done, hist, fin = handle.simulate_rollout_timed(pol, opt, 3, steps)
    assert rc == 0 and done == steps
    # annotation set at t=0, timeout exceeded at the first reconcile with now > 100 (t=120, reconcile 4): pod deletion is
    # not enabled, so the node goes on from pod-deletion-required to drain-required and further
    assert [abi.STATE_NAMES[c & 15] for c in fin["state"]] == state
    assert all(((fin["flags"][i] & abi.UST_F_WAIT_START_ANNO) != 0) == (start[i] is not None) for i in range(3))
    assert hist[4]["hist"][abi.UST_STATE_WAIT_FOR_JOBS_REQUIRED] == 3 and hist[5]["hist"][abi.UST_STATE_WAIT_FOR_JOBS_REQUIRED] == 0


@pytest.mark.parametrize("n", [0, 1, 5000, 700_001])
def test_packed_host_format(handle, n):
    """ust_apply_state_packed (uint16 revisions, int8 DaemonSet indices over PCIe, widened on the device) ==
    ust_apply_state == the oracle, on the direct and the pipelined upload path; the snapshot it leaves resident
    takes delta updates."""
    rng = np.random.default_rng(77 + n)
    soa, _ = helpers.random_soa(rng, n, wild=True, p_err=1e-4 if n == 5000 else 0.0)
    soa["pod_rev"] = rng.integers(0, 65536, n).astype(np.int32)      # the whole uint16 range
    soa["ds_rev"] = rng.integers(0, 65536, 3).astype(np.int32)
    soa["pod_rev"][::3] = soa["ds_rev"][rng.integers(0, 3, soa["pod_rev"][::3].shape[0])]
    for rep in range(2):
        pol = helpers.random_policy(rng)
        got = handle.apply_state_packed(pol, soa)
        ref = helpers.oracle_apply(pol, soa, variant=1)
        helpers.assert_same(got, ref, f"packed n={n} rep={rep}")
        helpers.assert_same(gpu_apply(handle, pol, soa), ref, f"wide n={n} rep={rep}")
    if n >= 5000:
        handle.apply_state_packed(pol, soa)
        idx = rng.choice(n, size=n // 50, replace=False).astype(np.int64)
        fresh, _ = helpers.random_soa(rng, idx.shape[0], wild=True)
        for k in ("state", "flags", "pod_rev", "ds_idx"):
            soa[k][idx] = fresh[k]
        got = handle.apply_state_delta(pol, n, idx, {k: fresh[k] for k in ("state", "flags", "pod_rev", "ds_idx")}, soa["ds_rev"])
        helpers.assert_same(got, helpers.oracle_apply(pol, soa, variant=1), "delta after a packed call")
    too_many = dict(soa); too_many["ds_rev"] = np.zeros(128, np.int32)
    assert handle.apply_state_packed(pol, too_many)[0] == abi.K["UST_ERR_INVALID_ARGUMENT"]


def test_many_daemonsets_use_the_global_table(handle):
    rng = np.random.default_rng(99)
    n, n_ds = 50_000, 3000  # > UST_DS_SMEM_MAX
    soa, _ = helpers.random_soa(rng, n, n_ds=n_ds)
    pol = abi.make_policy(max_parallel_upgrades=7, max_unavailable="25%")
    helpers.assert_same(gpu_apply(handle, pol, soa), helpers.oracle_apply(pol, soa, variant=0), "n_ds=3000")


def test_disabled_and_nil_policy(handle):
    rng = np.random.default_rng(5)
    soa, _ = helpers.random_soa(rng, 10_000, p_err=0.01)
    off = abi.make_policy(auto_upgrade=False, max_parallel_upgrades=3)
    for nil in (False, True):
        got = gpu_apply(handle, off, soa, nil_policy=nil)
        ref = helpers.oracle_apply(off, soa, variant=0, nil_policy=nil)
        helpers.assert_same(got, ref, f"nil={nil}")
        assert got[0] == 0 and not got[2].any()


def test_idempotent_and_stateless(handle):
    """ApplyState is stateless (upgrade_state.go:166-170): same snapshot, same answer, whatever ran before."""
    rng = np.random.default_rng(17)
    soa, pods = helpers.random_soa(rng, 40_000, with_pods=True)
    pol = helpers.random_policy(rng)
    a = gpu_apply(handle, pol, soa, pods)
    soa2, _ = helpers.random_soa(rng, 1234, p_err=0.05)
    gpu_apply(handle, helpers.random_policy(rng), soa2)  # an aborting call in between
    b = gpu_apply(handle, pol, soa, pods)
    helpers.assert_same(a, b, "repeat")


# ---- BASELINE.json configurations at full size ------------------------------------------------------

@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_baseline_config_bit_exact(handle, name):
    cfg = synth.CONFIGS[name]
    soa = synth.make_nodes(cfg["n"], cfg["seed"])
    pol = synth.config_policy(name)
    got = gpu_apply(handle, pol, soa)
    ref = helpers.oracle_apply(pol, soa, variant=1)  # SoA oracle: seconds at 10 M nodes
    helpers.assert_same(got, ref, name)
    if cfg["n"] <= 1_000_000:
        helpers.assert_same(got, helpers.oracle_apply(pol, soa, variant=0), name + " (reference-shaped)")
    rc, nxt, act, oc, cnt = got
    # size-independent properties
    code = soa["state"] & 15
    assert cnt["hist"][:14] == [int(np.sum(code == c)) for c in range(14)]
    granted = int(np.sum((code == 1) & (nxt == 2) & ((soa["state"] & abi.UST_HOT_UNSCHEDULABLE) == 0)))
    assert granted <= max(cnt["upgrades_available"], 0)
    assert np.all((act & abi.UST_A_SET_STATE != 0) == (nxt != code))


def test_c4_pod_lists_sample(handle):
    """C4 (pod-list actuators) on a 300 k-node sample against the reference-shaped oracle."""
    cfg = synth.CONFIGS["C4"]
    n = 300_000
    soa = synth.make_nodes(n, cfg["seed"])
    pods = synth.make_pods(n, cfg["seed"])
    pol = synth.config_policy("C4")
    helpers.assert_same(gpu_apply(handle, pol, soa, pods), helpers.oracle_apply(pol, soa, pods, variant=0), "C4 sample")


def test_c4_full_size(handle):
    """C4 at BASELINE.json's size: 10 M nodes with ~300 M workload pods in CSR lists (pod deletion and drain
    enabled). Bit-exact against the SoA oracle; the SoA oracle itself is checked against the reference-shaped
    one (an independent restatement of rows 12-14) on a 200 k-node slice of the same input."""
    cfg = synth.CONFIGS["C4"]
    n = cfg["n"]
    soa = synth.make_nodes(n, cfg["seed"])
    pods = synth.make_pods_blocked(n, cfg["seed"])
    pol = synth.config_policy("C4")
    got = gpu_apply(handle, pol, soa, pods)
    ref = helpers.oracle_apply(pol, soa, pods, variant=1)
    helpers.assert_same(got, ref, "C4 full size")
    rc, nxt, act, oc, cnt = got
    code = soa["state"] & 15
    assert rc == 0 and cnt["total_managed"] == int(np.isin(code, [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12]).sum())
    # reference-shaped oracle on the first 200 k nodes (independent restatement of rows 12-14)
    m = 200_000
    sub = {k: (v[:m] if k != "ds_rev" else v) for k, v in soa.items()}
    psub = {"pod_off": pods["pod_off"][:m + 1].copy(), "pod_flags": pods["pod_flags"][:int(pods["pod_off"][m])]}
    a = helpers.oracle_apply(pol, sub, psub, variant=0)
    b = helpers.oracle_apply(pol, sub, psub, variant=1)
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[1], b[1])


def test_device_resident_entry_point(handle):
    """ust_apply_state_device on torch-owned device buffers (plumbing only) == host entry point."""
    import torch
    cfg = synth.CONFIGS["C2"]
    soa = synth.make_nodes(cfg["n"], cfg["seed"])
    pol = synth.config_policy("C2")
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(v).to(dev) for k, v in soa.items()}
    n = cfg["n"]
    nxt = torch.empty(n, dtype=torch.uint8, device=dev)
    act = torch.empty(n, dtype=torch.int16, device=dev)
    cnt = torch.zeros(C.sizeof(abi.Counters) // 8, dtype=torch.int64, device=dev)
    handle.apply_state_device(pol, n, t["state"].data_ptr(), t["flags"].data_ptr(), t["pod_rev"].data_ptr(),
                              t["ds_idx"].data_ptr(), 4, t["ds_rev"].data_ptr(), nxt.data_ptr(), act.data_ptr(),
                              counters=cnt.data_ptr())
    handle.sync()
    ref = helpers.oracle_apply(pol, soa, variant=1)
    assert np.array_equal(nxt.cpu().numpy(), ref[1])
    assert np.array_equal(act.cpu().numpy().view(np.uint16), ref[2])
    c = abi.Counters.from_buffer_copy(cnt.cpu().numpy().tobytes())
    assert c.as_dict() == ref[4]


@pytest.mark.parametrize("shared_outputs", [False, True])
def test_back_to_back_device_calls_overlap_or_not(handle, shared_outputs):
    """Device-resident calls enqueued back to back on the handle's own stream, without a sync in between. With separate
    buffers per call they overlap (the next call streams while the previous one is still being decided, its tiles redone,
    its abort masked); with shared output arrays they run in strict order. Either way every call's outputs and counters
    are the oracle's."""
    import torch
    dev = torch.device("cuda:0")
    n = 1_200_037
    sets = 3
    rng = np.random.default_rng(99)
    snaps, bufs = [], []
    for k in range(sets):
        soa, _ = helpers.random_soa(rng, n, all_states=True, wild=True, p_err=(0.0, 0.0, 2e-6)[k])
        snaps.append(soa)
        t = {key: torch.from_numpy(v).to(dev) for key, v in soa.items()}
        t["next"] = torch.empty(n, dtype=torch.uint8, device=dev)
        t["actions"] = torch.empty(n, dtype=torch.int16, device=dev)
        t["cnt"] = torch.zeros(C.sizeof(abi.Counters) // 8, dtype=torch.int64, device=dev)
        bufs.append(t)
    if shared_outputs:
        for t in bufs[1:]:
            t["next"], t["actions"] = bufs[0]["next"], bufs[0]["actions"]
    pols = [abi.make_policy(max_parallel_upgrades=0, max_unavailable="30%"), abi.make_policy(max_parallel_upgrades=100, max_unavailable="25%"),
            abi.make_policy(max_parallel_upgrades=0), abi.make_policy(max_parallel_upgrades=n // 9, max_unavailable="60%")]
    calls = 14
    for c in range(calls):
        t, pol = bufs[c % sets], pols[c % len(pols)]
        handle.apply_state_device(pol, n, t["state"].data_ptr(), t["flags"].data_ptr(), t["pod_rev"].data_ptr(), t["ds_idx"].data_ptr(),
                                  len(snaps[c % sets]["ds_rev"]), t["ds_rev"].data_ptr(), t["next"].data_ptr(), t["actions"].data_ptr(),
                                  counters=t["cnt"].data_ptr())
    handle.sync()
    last = {}
    for c in range(calls):
        last[c % sets] = c
    for k, c in last.items():
        if shared_outputs and c != calls - 1:
            continue   # the shared arrays hold the last call's outputs
        ref = helpers.oracle_apply(pols[c % len(pols)], snaps[k], variant=1)
        assert np.array_equal(bufs[k]["next"].cpu().numpy(), ref[1]), (k, c)
        assert np.array_equal(bufs[k]["actions"].cpu().numpy().view(np.uint16), ref[2]), (k, c)
        assert abi.Counters.from_buffer_copy(bufs[k]["cnt"].cpu().numpy().tobytes()).as_dict() == ref[4], (k, c)


def test_overlapped_calls_every_call_checked(handle):
    """24 device calls enqueued without a sync, each with output arrays of its own, over four snapshots (one of them
    aborting) and six policies: calls k and k+1 share nothing but read-only inputs, so k+1 streams while k is decided,
    its tiles redone or its abort masked - and EVERY call's outputs and counters must be the oracle's, not only the
    last ones. The handle must report that calls did overlap (else this test checks nothing new)."""
    import torch
    dev = torch.device("cuda:0")
    n = 700_019
    rng = np.random.default_rng(4242)
    snaps, ins = [], []
    for k in range(4):
        soa, _ = helpers.random_soa(rng, n, all_states=True, wild=(k % 2 == 1), p_err=(0.0, 0.0, 0.0, 3e-6)[k])
        snaps.append(soa)
        ins.append({key: torch.from_numpy(v).to(dev) for key, v in soa.items()})
    pols = [abi.make_policy(max_parallel_upgrades=0, max_unavailable="30%"), abi.make_policy(max_parallel_upgrades=0, max_unavailable="30%"),
            abi.make_policy(max_parallel_upgrades=100, max_unavailable="25%"), abi.make_policy(max_parallel_upgrades=0),
            abi.make_policy(max_parallel_upgrades=n // 7, max_unavailable="55%"), abi.make_policy(max_parallel_upgrades=0, max_unavailable="12%")]
    calls = 24
    outs = [{"next": torch.empty(n, dtype=torch.uint8, device=dev), "actions": torch.empty(n, dtype=torch.int16, device=dev),
             "cnt": torch.zeros(C.sizeof(abi.Counters) // 8, dtype=torch.int64, device=dev)} for _ in range(calls)]
    torch.cuda.synchronize()
    before = handle.overlapped_calls()
    for c in range(calls):
        t, o = ins[c % 4], outs[c]
        handle.apply_state_device(pols[c % len(pols)], n, t["state"].data_ptr(), t["flags"].data_ptr(), t["pod_rev"].data_ptr(),
                                  t["ds_idx"].data_ptr(), len(snaps[c % 4]["ds_rev"]), t["ds_rev"].data_ptr(), o["next"].data_ptr(),
                                  o["actions"].data_ptr(), counters=o["cnt"].data_ptr())
    handle.sync()
    assert handle.overlapped_calls() - before >= calls // 2, "the calls did not overlap: nothing was tested"
    refs = {}
    for c in range(calls):
        key = (c % len(pols), c % 4)
        if key not in refs:
            refs[key] = helpers.oracle_apply(pols[key[0]], snaps[key[1]], variant=1)
        ref = refs[key]
        assert np.array_equal(outs[c]["next"].cpu().numpy(), ref[1]), c
        assert np.array_equal(outs[c]["actions"].cpu().numpy().view(np.uint16), ref[2]), c
        assert abi.Counters.from_buffer_copy(outs[c]["cnt"].cpu().numpy().tobytes()).as_dict() == ref[4], c


def test_device_entry_point_with_moving_budget_cut(handle):
    """One fused launch over 2.5 M device-resident nodes, the slot budget cutting through the middle of the
    array; consecutive calls share the size and policy (so the second and later ones run on the previous call's
    cut), the snapshots differ."""
    import torch
    dev = torch.device("cuda:0")
    n = 2_500_063
    pol = abi.make_policy(max_parallel_upgrades=0, max_unavailable="30%")
    nxt = torch.empty(n, dtype=torch.uint8, device=dev)
    act = torch.empty(n, dtype=torch.int16, device=dev)
    cnt = torch.zeros(C.sizeof(abi.Counters) // 8, dtype=torch.int64, device=dev)
    for rep, frac in enumerate((0.6, 0.6, 0.9, 0.32, 0.6)):
        rng = np.random.default_rng(5000 + (rep if rep != 1 else 0))
        soa, _ = helpers.random_soa(rng, n, all_states=False, wild=False)
        pick = rng.random(n)
        soa["state"] = np.where(pick < frac, (soa["state"] & 0xF0) | 1, soa["state"]).astype(np.uint8)
        soa["state"] &= np.uint8(0x7F)
        t = {k: torch.from_numpy(v).to(dev) for k, v in soa.items()}
        handle.apply_state_device(pol, n, t["state"].data_ptr(), t["flags"].data_ptr(), t["pod_rev"].data_ptr(),
                                  t["ds_idx"].data_ptr(), len(soa["ds_rev"]), t["ds_rev"].data_ptr(), nxt.data_ptr(),
                                  act.data_ptr(), counters=cnt.data_ptr())
        handle.sync()
        ref = helpers.oracle_apply(pol, soa, variant=1)
        assert np.array_equal(nxt.cpu().numpy(), ref[1]), f"rep={rep}"
        assert np.array_equal(act.cpu().numpy().view(np.uint16), ref[2]), f"rep={rep}"
        c = abi.Counters.from_buffer_copy(cnt.cpu().numpy().tobytes())
        assert c.as_dict() == ref[4], f"rep={rep}"
