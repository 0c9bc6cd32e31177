"""The oracle against every known-answer test the reference holds for the path."""
import ctypes as C

import numpy as np
import pytest

import helpers
from helpers import abi

G = helpers.load_golden()
VEC = [v for v in G["apply_state"]]


def _run(v, variant):
    pol = helpers.policy_from_vector(v)
    soa, pods = helpers.encode_nodes(v["nodes"], G["daemonset_hash"], v.get("policy"))
    if pol is None:
        pol_arg, nil = abi.Policy(), True
    else:
        pol_arg, nil = pol, False
    return soa, helpers.oracle_apply(pol_arg, soa, pods, variant=variant, nil_policy=nil)


@pytest.mark.parametrize("variant", [0, 1], ids=["reference-shaped", "soa-scalar"])
@pytest.mark.parametrize("v", VEC, ids=[f'{v["ref"].split("/")[-1]}' for v in VEC])
def test_apply_state_vector(v, variant):
    if v.get("nil_state"):
        # upgrade_state.go:175-177 — at the ABI a nil snapshot is NULL arrays
        rc = helpers.oracle().ust_oracle_apply_state(C.c_int(variant), None, C.c_int64(1), None, None, None, None,
                                                     C.c_int32(0), None, None, None, None, None, None)
        assert rc != 0
        return
    soa, (rc, nxt, act, out, cnt) = _run(v, variant)
    n = helpers.check_vector(v, soa, rc, nxt, act, out)
    assert n > 0 or v.get("actuator_error"), "vector asserts nothing"
    helpers.check_derived(v, nxt)


def test_vector_count_matches_survey():
    # SURVEY.md §8c lists the known-answer tests; make sure none was dropped from the fixture
    refs = {v["ref"] for v in VEC}
    assert len(VEC) >= 55
    for must in (":196-225", ":413-440", ":441-513", ":972-1017", ":1268-1294", ":1566-1609"):
        assert any(r.endswith(must) for r in refs), must


@pytest.mark.parametrize("s", G["intstr"], ids=lambda s: f'{s["percent"]}pct_of_{s["total"]}')
def test_intstr_scaling(s):
    out = C.c_int64()
    assert helpers.oracle().ust_oracle_scaled_value(1, C.c_int64(s["percent"]), C.c_int64(s["total"]), C.byref(out)) == 0
    assert out.value == s["expect"]


@pytest.mark.parametrize("b", G["build_state"], ids=lambda b: b["name"][:40])
def test_build_state_vector(b):
    pods = b["pods"]
    n = len(pods)
    state = np.zeros(n, np.uint8)
    ds_idx = np.full(n, -1, np.int32)
    for i, p in enumerate(pods):
        code = abi.STATE_CODE.get(p["node_state"], abi.UST_STATE_OTHER)
        if p["node_name"] == "" and p["phase"] == "Pending":  # upgrade_state.go:149-152
            code = abi.UST_STATE_EXCLUDED
        state[i] = code
        ds_idx[i] = -1 if p["ds"] is None else p["ds"]
    desired = np.array([d["desired"] for d in b["daemonsets"]], np.int32)
    cnt = abi.Counters()
    rc = helpers.oracle().ust_oracle_build_state(
        C.c_int64(n), state.ctypes.data_as(C.c_void_p), ds_idx.ctypes.data_as(C.c_void_p),
        C.c_int32(len(desired)), desired.ctypes.data_as(C.c_void_p), C.byref(cnt))
    if b["expect_error"]:
        assert rc == abi.K["UST_ERR_" + b["expect_error"]]
        return
    assert rc == 0
    got = {abi.STATE_NAMES[c]: cnt.hist[c] for c in range(13) if cnt.hist[c]}
    assert got == b["expect_buckets"]


# ---- BuildState with the owner join at UID level (the device path's wire format) -------------------

@pytest.mark.parametrize("b", G["build_state"], ids=lambda b: b["name"][:40])
def test_build_state_vector_uid_join(b):
    """The reference's BuildState specs (upgrade_state_test.go:122-185) restated with owner UIDs instead of
    pre-resolved DaemonSet indices: same buckets, same error, and the join returns the indices."""
    rng = np.random.default_rng(11)
    state, owner, ds_uid, desired = helpers.uid_inputs_from_vector(b, rng)
    rc, ds_idx, cnt = helpers.oracle_build_state_uids(state, owner, ds_uid, desired)
    if b["expect_error"]:
        assert rc == abi.K["UST_ERR_" + b["expect_error"]]
        return
    assert rc == 0
    got = {abi.STATE_NAMES[c]: cnt["hist"][c] for c in range(13) if cnt["hist"][c]}
    assert got == b["expect_buckets"]
    assert list(ds_idx) == [(-1 if p["ds"] is None else p["ds"]) for p in b["pods"]]
    for i in b.get("expect_orphan", []):
        assert ds_idx[i] == -1


def test_build_state_uid_join_equals_index_form():
    """UID form == index form once foreign-owned pods (dropped by GetPodsOwnedbyDs / GetOrphanedPods,
    common_manager.go:190-222) are taken out of the index form's input."""
    rng = np.random.default_rng(5)
    for n, n_ds in ((0, 0), (1, 1), (1000, 3), (5000, 40), (20_000, 1500)):
        ds_uid = rng.integers(1, 2 ** 63, size=(n_ds, 2), dtype=np.uint64)
        kind = rng.random(n)
        truth = np.where(kind < 0.05, -1, np.where(kind < 0.15, -2, rng.integers(0, max(n_ds, 1), n))).astype(np.int32)
        if n_ds == 0:
            truth[truth >= 0] = -1
        owner = np.zeros((n, 2), np.uint64)
        owner[truth >= 0] = ds_uid[truth[truth >= 0]] if n_ds else 0
        foreign = truth == -2
        owner[foreign] = rng.integers(2 ** 63, 2 ** 64 - 1, size=(int(foreign.sum()), 2), dtype=np.uint64)  # never a DaemonSet's
        state = rng.integers(0, 16, n).astype(np.uint8) | (rng.integers(0, 16, n).astype(np.uint8) << 4)
        desired = np.bincount(truth[truth >= 0], minlength=n_ds).astype(np.int32)
        if n_ds > 2:
            desired[n_ds // 2] += int(rng.integers(0, 2))  # sometimes a DaemonSet with unscheduled pods
        rc, ds_idx, cnt = helpers.oracle_build_state_uids(state, owner, ds_uid, desired)
        assert np.array_equal(ds_idx, truth)
        keep = ~foreign
        ref = abi.Counters()
        st = np.ascontiguousarray(state[keep]); di = np.ascontiguousarray(truth[keep])
        rc2 = helpers.oracle().ust_oracle_build_state(
            C.c_int64(int(keep.sum())), st.ctypes.data_as(C.c_void_p), di.ctypes.data_as(C.c_void_p),
            C.c_int32(n_ds), desired.ctypes.data_as(C.c_void_p), C.byref(ref))
        assert rc == rc2
        r = ref.as_dict()
        r["hist"] = list(r["hist"]); r["hist"][14] += int(foreign.sum())   # dropped pods show up as "not in snapshot"
        c = dict(cnt); c["hist"] = list(c["hist"])
        assert c == r
    # duplicate / empty DaemonSet UIDs are rejected
    dup = np.array([[1, 2], [1, 2]], np.uint64)
    assert helpers.oracle_build_state_uids(np.zeros(0, np.uint8), np.zeros((0, 2), np.uint64), dup, np.zeros(2, np.int32))[0] == abi.K["UST_ERR_INVALID_ARGUMENT"]
