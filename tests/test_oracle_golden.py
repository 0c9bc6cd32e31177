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
