"""Test infrastructure: oracle loader, golden-vector encoder, comparison helpers."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "k8s-operator-libs_b200"))

from ust import abi  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
ORACLE_SO = os.path.join(ROOT, "oracle", "libust_oracle.so")

_oracle = None


def oracle():
    """The CPU oracle (oracle/ust_oracle.cpp). Test-only."""
    global _oracle
    if _oracle is None:
        src = os.path.join(ROOT, "oracle", "ust_oracle.cpp")
        newest = max(os.path.getmtime(src), os.path.getmtime(abi.HEADER))
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < newest:
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        lib = C.CDLL(ORACLE_SO)
        lib.ust_oracle_apply_state.restype = C.c_int
        lib.ust_oracle_time_apply_state.restype = C.c_double
        lib.ust_oracle_scaled_value.restype = C.c_int
        lib.ust_oracle_build_state.restype = C.c_int
        lib.ust_oracle_build_state_uids.restype = C.c_int
        lib.ust_oracle_simulate.restype = C.c_int
        _oracle = lib
    return _oracle


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pods_struct(pods):
    if pods is None:
        return None, None
    off = np.ascontiguousarray(pods["pod_off"], dtype=np.int32)
    pf = np.ascontiguousarray(pods["pod_flags"], dtype=np.uint16)
    s = abi.Pods(off.ctypes.data, pf.ctypes.data, int(pf.shape[0]))
    return s, (off, pf)


def oracle_apply(policy, soa, pods=None, variant=0, nil_policy=False):
    """Run the oracle on SoA arrays. Returns (rc, next_state, actions, outcome, counters-dict)."""
    n = int(soa["state"].shape[0])
    nxt = np.zeros(n, np.uint8)
    act = np.zeros(n, np.uint16)
    out = np.full(n, 0xFF, np.uint8)
    cnt = abi.Counters()
    ps, keep = pods_struct(pods)
    rc = oracle().ust_oracle_apply_state(
        C.c_int(variant), None if nil_policy else C.byref(policy), C.c_int64(n), _ptr(soa["state"]), _ptr(soa["flags"]),
        _ptr(soa["pod_rev"]), _ptr(soa["ds_idx"]), C.c_int32(int(soa["ds_rev"].shape[0])), _ptr(soa["ds_rev"]),
        C.byref(ps) if ps is not None else None, _ptr(nxt), _ptr(act), _ptr(out), C.byref(cnt))
    del keep
    return rc, nxt, act, out, cnt.as_dict()


# ---- golden vector -> SoA (the test-side mirror of the host encoder) --------------------------------

def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def policy_from_vector(v):
    p = v.get("policy")
    if p is None:
        return None
    opts = v.get("options") or {}
    return abi.make_policy(
        auto_upgrade=p.get("autoUpgrade", False),
        max_parallel_upgrades=p.get("maxParallelUpgrades", 0),
        max_unavailable=p.get("maxUnavailable"),
        pod_deletion_enabled=opts.get("podDeletionEnabled", False),
        validation_enabled=opts.get("validationEnabled", False),
        pod_deletion=p.get("podDeletion"),
        drain=p.get("drain"),
        wait_for_completion=p.get("waitForCompletion"),
        use_maintenance_operator=opts.get("useMaintenanceOperator", False),
        evaluate_actuators=v.get("evaluate_actuators", False),
    )


_PHASES = {"Pending": abi.UST_PHASE_PENDING, "Running": abi.UST_PHASE_RUNNING,
           "Succeeded": abi.UST_PHASE_SUCCEEDED, "Failed": abi.UST_PHASE_FAILED}


def encode_nodes(nodes, ds_hash, policy_dict=None):
    """Encode golden-vector node descriptions into the SoA arrays of include/ust.h.

    Each predicate is evaluated the way the reference evaluates it (citations in include/ust.h).
    """
    n = len(nodes)
    hot = np.zeros(n, np.uint8)
    flags = np.zeros(n, np.uint32)
    pod_rev = np.zeros(n, np.int32)
    ds_idx = np.full(n, -1, np.int32)
    intern = {ds_hash: 1}
    pod_off = [0]
    pod_flags = []
    any_workload = any("workload" in nd for nd in nodes)
    timeout = ((policy_dict or {}).get("waitForCompletion") or {}).get("timeoutSeconds", 0)
    for i, nd in enumerate(nodes):
        code = abi.STATE_CODE.get(nd.get("state", ""), abi.UST_STATE_OTHER)
        h = code
        if nd.get("unschedulable"):
            h |= abi.UST_HOT_UNSCHEDULABLE
        if nd.get("ready") not in (None, "True"):
            h |= abi.UST_HOT_NOT_READY
        if nd.get("skip") == "true":
            h |= abi.UST_HOT_SKIP
        f = 0
        anno = nd.get("anno") or {}
        if anno.get("upgrade-requested") == "true":
            f |= abi.UST_F_UPGRADE_REQUESTED
        if anno.get("safe-load", "") != "":
            f |= abi.UST_F_SAFE_LOAD
        if "initial-state" in anno:
            f |= abi.UST_F_INITIAL_STATE_ANNO
        if "requestor-mode" in anno:
            f |= abi.UST_F_REQUESTOR_MODE
        if "wait-start" in anno:
            f |= abi.UST_F_WAIT_START_ANNO
            ws = anno["wait-start"]
            if ws.startswith("now-"):
                if int(ws[4:]) > timeout:  # currentTime > startTime + timeoutSeconds  pod_manager.go:354
                    f |= abi.UST_F_WAIT_TIMED_OUT
            else:
                f |= abi.UST_F_WAIT_START_INVALID
        if nd.get("validation_done", True):
            f |= abi.UST_F_VALIDATION_DONE
        has_ds = bool(nd.get("ds"))
        pod = nd.get("pod")
        if not has_ds:
            f |= abi.UST_F_POD_ORPHANED
        else:
            ds_idx[i] = 0
            if pod is None or "hash" not in pod:
                h |= abi.UST_HOT_REVISION_HASH_ERROR  # pod_manager.go:84-89
            else:
                pod_rev[i] = intern.setdefault(pod["hash"], len(intern) + 1)
        if pod:
            ctrs = pod.get("containers", [])
            init = pod.get("init", [])
            if pod.get("phase") == "Running" and len(ctrs) != 0 and all(c[0] for c in ctrs):
                f |= abi.UST_F_POD_READY
            if any((not c[0]) and c[1] > 10 for c in init + ctrs):
                f |= abi.UST_F_POD_FAILING
            if pod.get("terminating"):
                f |= abi.UST_F_POD_TERMINATING
        nm = nd.get("nm")
        if nm is not None:
            f |= abi.UST_F_NM_PRESENT
            if nm.get("ready"):
                f |= abi.UST_F_NM_READY
        wl = nd.get("workload", [])
        for wp in wl:
            pf = _PHASES.get(wp.get("phase"), abi.UST_PHASE_OTHER)
            if wp.get("controller"):
                pf |= abi.UST_POD_HAS_CONTROLLER
                if wp["controller"] == "DaemonSet":
                    pf |= abi.UST_POD_CONTROLLED_BY_DS
            if wp.get("ds_missing"):
                pf |= abi.UST_POD_DS_MISSING
            if wp.get("mirror"):
                pf |= abi.UST_POD_MIRROR
            if wp.get("emptydir"):
                pf |= abi.UST_POD_HAS_EMPTYDIR
            if wp.get("match_filter"):
                pf |= abi.UST_POD_MATCH_DELETION_FILTER
            if wp.get("match_wait"):
                pf |= abi.UST_POD_MATCH_WAIT_SELECTOR
            if wp.get("match_drain", True):
                pf |= abi.UST_POD_MATCH_DRAIN_SELECTOR
            pod_flags.append(pf)
        if any(wp.get("match_wait") and wp.get("phase") in ("Running", "Pending") for wp in wl):
            f |= abi.UST_F_WAIT_PODS_RUNNING
        pod_off.append(len(pod_flags))
        hot[i] = h
        flags[i] = f
    soa = {"state": hot, "flags": flags, "pod_rev": pod_rev, "ds_idx": ds_idx,
           "ds_rev": np.array([1], np.int32)}
    pods = None
    if any_workload:
        pods = {"pod_off": np.array(pod_off, np.int32), "pod_flags": np.array(pod_flags, np.uint16)}
    return soa, pods


def check_vector(v, soa, rc, nxt, act, outcome):
    """Assert exactly what the Go test asserts (see make_reference_vectors.py). Returns #assertions."""
    checks = 0
    nodes = v["nodes"]
    names = [abi.STATE_NAMES[c] if c < 13 else "other" for c in nxt]
    err = v.get("expect_error", None)
    if "expect_error" in v and v.get("actuator_error") is None:
        if err is None:
            assert rc == 0, (v["name"], rc)
        else:
            assert rc == abi.K["UST_ERR_" + err], (v["name"], rc)
        checks += 1
    for i, nd in enumerate(nodes):
        ex = nd.get("expect")
        if not ex or v.get("actuator_error"):
            continue
        if "state" in ex:
            assert names[i] == ex["state"], (v["name"], i, names[i], ex["state"])
            checks += 1
        if "outcome" in ex:
            assert outcome[i] == abi.STATE_CODE[ex["outcome"]], (v["name"], i, outcome[i])
            checks += 1
        for a in ex.get("actions_present", []):
            assert act[i] & abi.ACTION_NAMES[a], (v["name"], i, a, hex(act[i]))
            checks += 1
        for a in ex.get("actions_absent", []):
            assert not (act[i] & abi.ACTION_NAMES[a]), (v["name"], i, a, hex(act[i]))
            checks += 1
        # annotation expectations, evaluated through the action bits + input flags
        fl = int(soa["flags"][i])
        present_after = {
            "initial-state": (bool(fl & abi.UST_F_INITIAL_STATE_ANNO) or bool(act[i] & abi.UST_A_SET_INITIAL_STATE_ANNO))
            and not (act[i] & abi.UST_A_CLEAR_INITIAL_STATE_ANNO),
            "upgrade-requested": bool(fl & abi.UST_F_UPGRADE_REQUESTED) and not (act[i] & abi.UST_A_CLEAR_UPGRADE_REQUESTED),
            "safe-load": bool(fl & abi.UST_F_SAFE_LOAD) and not (act[i] & abi.UST_A_UNBLOCK_SAFE_LOAD),
        }
        code_in = int(soa["state"][i]) & 15
        if act[i] & abi.UST_A_REQUESTOR_ANNO_CHANGE:  # set in upgrade-required, cleared in uncordon-required
            present_after["requestor-mode"] = code_in == abi.UST_STATE_UPGRADE_REQUIRED
        else:
            present_after["requestor-mode"] = bool(fl & abi.UST_F_REQUESTOR_MODE)
        for k in ex.get("anno_present", []):
            assert present_after[k], (v["name"], i, k)
            checks += 1
        for k in ex.get("anno_absent", []):
            assert not present_after[k], (v["name"], i, k)
            checks += 1
    if "expect_counts" in v:
        for st, c in v["expect_counts"].items():
            assert names.count(st) == c, (v["name"], st, names)
            checks += 1
    for group, total in v.get("expect_count_sums", []):
        assert sum(names.count(s) for s in group) == total, (v["name"], group, names)
        checks += 1
    for key, bit in (("expect_restart", abi.UST_A_RESTART_DRIVER_POD), ("expect_drain", abi.UST_A_SCHEDULE_DRAIN),
                     ("expect_eviction", abi.UST_A_SCHEDULE_POD_EVICTION), ("expect_uncordon", abi.UST_A_UNCORDON),
                     ("expect_nm_change", abi.UST_A_NM_CREATE_OR_DELETE)):
        if key in v and not v.get("actuator_error"):
            got = [i for i in range(len(nodes)) if act[i] & bit]
            assert got == v[key], (v["name"], key, got)
            checks += 1
    return checks


def check_derived(v, nxt):
    if "derived" in v:
        names = [abi.STATE_NAMES[c] if c < 13 else "other" for c in nxt]
        assert names == v["derived"], (v["name"], names)


def assert_same(a, b, what):
    """Bit-exact comparison of two (rc, next, actions, outcome, counters) results."""
    rc_a, n_a, a_a, o_a, c_a = a
    rc_b, n_b, a_b, o_b, c_b = b
    assert rc_a == rc_b, (what, "rc", rc_a, rc_b)
    bad = np.nonzero(n_a != n_b)[0]
    assert bad.size == 0, (what, "next_state", bad[:10], n_a[bad[:10]], n_b[bad[:10]])
    bad = np.nonzero(a_a != a_b)[0]
    assert bad.size == 0, (what, "actions", bad[:10], a_a[bad[:10]], a_b[bad[:10]])
    if o_a is not None and o_b is not None:
        bad = np.nonzero(o_a != o_b)[0]
        assert bad.size == 0, (what, "outcome", bad[:10], o_a[bad[:10]], o_b[bad[:10]])
    assert c_a == c_b, (what, "counters", c_a, c_b)


# ---- random inputs for cross-checks ------------------------------------------------------------------

def random_soa(rng, n, n_ds=3, p_err=0.0, with_pods=False, all_states=True, wild=True):
    """Uniformly random SoA (every flag bit i.i.d.), optionally with contract-edge cases (`wild`)."""
    codes = rng.integers(0, 16 if all_states else 13, size=n).astype(np.uint8)
    hot = codes.copy()
    for bit, p in ((abi.UST_HOT_NOT_READY, 0.2), (abi.UST_HOT_SKIP, 0.2), (abi.UST_HOT_UNSCHEDULABLE, 0.3),
                   (abi.UST_HOT_REVISION_HASH_ERROR, p_err)):
        hot |= np.where(rng.random(n) < p, bit, 0).astype(np.uint8)
    flags = np.zeros(n, np.uint32)
    for k, v in abi.K.items():
        if k.startswith("UST_F_") and k != "UST_F_INPUT_MASK":
            p = 0.15 if k == "UST_F_POD_ORPHANED" else 0.5
            flags |= np.where(rng.random(n) < p, np.uint32(v), np.uint32(0))
    if wild:
        flags |= (rng.integers(0, 2, size=n).astype(np.uint32) * np.uint32(0xFFC0061F))  # reserved bits set: must be ignored
    ds_rev = rng.integers(1, 4, size=n_ds).astype(np.int32)
    ds_idx = rng.integers(0, n_ds, size=n).astype(np.int32)
    if wild:
        ds_idx = np.where(rng.random(n) < 0.05, rng.integers(-3, n_ds + 3, size=n), ds_idx).astype(np.int32)
    orphan = (flags & abi.UST_F_POD_ORPHANED) != 0
    ds_idx = np.where(orphan & (rng.random(n) < 0.8), -1, ds_idx).astype(np.int32)
    pod_rev = rng.integers(1, 4, size=n).astype(np.int32)
    soa = {"state": hot, "flags": flags, "pod_rev": pod_rev, "ds_idx": ds_idx, "ds_rev": ds_rev}
    pods = None
    if with_pods:
        cnt = rng.integers(0, 7, size=n)
        off = np.zeros(n + 1, np.int32)
        np.cumsum(cnt, out=off[1:])
        total = int(off[-1])
        pf = rng.integers(0, 6, size=total).astype(np.uint16)  # phase incl. 0 and the undefined 5
        for k, v in abi.K.items():
            if k.startswith("UST_POD_") and k != "UST_POD_PHASE_MASK":
                pf |= np.where(rng.random(total) < 0.4, np.uint16(v), np.uint16(0))
        pods = {"pod_off": off, "pod_flags": pf}
    return soa, pods


def random_policy(rng, force_auto=True):
    mu = rng.choice(["nil", "int", "pct", "pct100", "bad"], p=[0.3, 0.25, 0.3, 0.1, 0.05])
    max_unav = {"nil": None, "int": int(rng.integers(0, 40)), "pct": f"{int(rng.integers(0, 101))}%",
                "pct100": "100%", "bad": "a-few"}[mu]
    pd_present = rng.random() < 0.8
    return abi.make_policy(
        auto_upgrade=True if force_auto else rng.random() < 0.9,
        max_parallel_upgrades=int(rng.choice([0, 0, 1, 3, 10, 100, 1000])),
        max_unavailable=max_unav,
        pod_deletion_enabled=rng.random() < 0.5,
        validation_enabled=rng.random() < 0.5,
        pod_deletion={"force": bool(rng.random() < 0.5), "deleteEmptyDir": bool(rng.random() < 0.5)} if pd_present else None,
        drain={"enable": bool(rng.random() < 0.6), "force": bool(rng.random() < 0.5),
               "deleteEmptyDir": bool(rng.random() < 0.5)} if rng.random() < 0.8 else None,
        wait_for_completion={"podSelector": "app=x" if rng.random() < 0.6 else "",
                             "timeoutSeconds": int(rng.choice([0, 30]))} if rng.random() < 0.7 else None,
        use_maintenance_operator=rng.random() < 0.3,
        evaluate_actuators=rng.random() < 0.7,
    )


def oracle_build_state_uids(state, owner_uid, ds_uid, ds_desired):
    """Reference-shaped BuildState with the owner join at UID level. Returns (rc, ds_idx, counters-dict)."""
    n = int(state.shape[0])
    owner_uid = np.ascontiguousarray(owner_uid, dtype=np.uint64).reshape(n, 2)
    ds_uid = np.ascontiguousarray(ds_uid, dtype=np.uint64).reshape(-1, 2)
    ds_desired = np.ascontiguousarray(ds_desired, dtype=np.int32)
    ds_idx = np.full(n, -3, np.int32)
    cnt = abi.Counters()
    rc = oracle().ust_oracle_build_state_uids(
        C.c_int64(n), _ptr(state), _ptr(owner_uid), C.c_int32(int(ds_uid.shape[0])), _ptr(ds_uid), _ptr(ds_desired),
        _ptr(ds_idx), C.byref(cnt))
    return rc, ds_idx, cnt.as_dict()


def uid_inputs_from_vector(b, rng):
    """BuildState golden vector -> UID-level inputs: every DaemonSet gets a random 128-bit UID, a pod carries its
    owner's UID ((0, 0) when the vector says it has no owner reference)."""
    pods = b["pods"]
    n = len(pods)
    ds_uid = rng.integers(1, 2 ** 63, size=(len(b["daemonsets"]), 2), dtype=np.uint64)
    state = np.zeros(n, np.uint8)
    owner = np.zeros((n, 2), np.uint64)
    for i, p in enumerate(pods):
        code = abi.STATE_CODE.get(p["node_state"], abi.UST_STATE_OTHER)
        if p["node_name"] == "" and p["phase"] == "Pending":  # upgrade_state.go:149-152
            code = abi.UST_STATE_EXCLUDED
        state[i] = code
        if p["ds"] is not None:
            owner[i] = ds_uid[p["ds"]]
    desired = np.array([d["desired"] for d in b["daemonsets"]], np.int32)
    return state, owner, ds_uid, desired


def oracle_simulate(policy, soa, steps, variant=1):
    """CPU rollout simulation (oracle ApplyState + the feedback restatement). Returns (rc, steps_done, history, final)."""
    n = int(soa["state"].shape[0])
    fin = {"state": soa["state"].copy(), "flags": soa["flags"].copy(), "pod_rev": soa["pod_rev"].copy()}
    hist = (abi.Counters * max(steps, 1))()
    done = C.c_int32(0)
    rc = oracle().ust_oracle_simulate(
        C.c_int(variant), C.byref(policy) if policy is not None else None, C.c_int64(n), _ptr(fin["state"]), _ptr(fin["flags"]),
        _ptr(fin["pod_rev"]), _ptr(soa["ds_idx"]), C.c_int32(int(soa["ds_rev"].shape[0])), _ptr(soa["ds_rev"]), C.c_int32(steps),
        hist, C.byref(done))
    return rc, int(done.value), [hist[k].as_dict() for k in range(steps)], fin


def oracle_simulate_timed(policy, options, soa, steps, variant=1):
    """CPU restatement of ust_simulate_rollout_timed (options None: the untimed feedback, incl. requestor mode)."""
    n = int(soa["state"].shape[0])
    fin = {"state": soa["state"].copy(), "flags": soa["flags"].copy(), "pod_rev": soa["pod_rev"].copy()}
    hist = (abi.Counters * max(steps, 1))()
    done = C.c_int32(0)
    rc = oracle().ust_oracle_simulate_timed(
        C.c_int(variant), C.byref(policy) if policy is not None else None, C.byref(options) if options is not None else None,
        C.c_int64(n), _ptr(fin["state"]), _ptr(fin["flags"]), _ptr(fin["pod_rev"]), _ptr(soa["ds_idx"]),
        C.c_int32(int(soa["ds_rev"].shape[0])), _ptr(soa["ds_rev"]), C.c_int32(steps), hist, C.byref(done))
    return rc, int(done.value), [hist[k].as_dict() for k in range(steps)], fin


def wait_timeout_timeline(golden_hash, timeout, dt, steps, n_nodes=3):
    """What the reference's HandleTimeoutOnPodCompletions (pod_manager.go:331-368) does to nodes whose wait-selector pods
    keep running, reconcile by reconcile at times k * dt: the annotation bookkeeping is done here, every reconcile is
    evaluated by the golden-vector encoder (the "wait-start": "now-N" rule of pod_manager_test.go:183-229) and the
    oracle's vector-pinned ApplyState. Returns (policy, initial soa, final state names, final start annotations)."""
    pol = abi.make_policy(max_parallel_upgrades=0, wait_for_completion={"podSelector": "app=job", "timeoutSeconds": timeout})
    pol.evaluate_actuators = 1
    nodes = [{"state": "wait-for-jobs-required", "ds": True, "pod": {"hash": golden_hash, "phase": "Running", "containers": [[True, 0]]}}
             for _ in range(n_nodes)]
    pdict = {"waitForCompletion": {"timeoutSeconds": timeout}}
    soa, _ = encode_nodes(nodes, golden_hash, pdict)
    soa["flags"] = (soa["flags"] | np.uint32(abi.UST_F_WAIT_PODS_RUNNING)).astype(np.uint32)   # the jobs keep running
    start = [None] * n_nodes
    state = ["wait-for-jobs-required"] * n_nodes
    for k in range(steps):
        now = k * dt
        vec = []
        for i in range(n_nodes):
            nd = dict(nodes[i], state=state[i])
            if start[i] is not None:
                nd["anno"] = {"wait-start": f"now-{now - start[i]}"}
            vec.append(nd)
        s2, _ = encode_nodes(vec, golden_hash, pdict)
        s2["flags"] = np.where((s2["state"] & 15) == 3, s2["flags"] | np.uint32(abi.UST_F_WAIT_PODS_RUNNING), s2["flags"]).astype(np.uint32)
        rc, nxt, act, oc, _ = oracle_apply(pol, s2, variant=0)
        assert rc == 0
        for i in range(n_nodes):
            if act[i] & abi.UST_A_SET_WAIT_START:
                start[i] = now
            if act[i] & abi.UST_A_CLEAR_WAIT_START:
                start[i] = None
            new = oc[i] if ((act[i] & abi.UST_A_SCHEDULE_WAIT_CHECK) and oc[i] != 0xFF) else nxt[i]
            state[i] = abi.STATE_NAMES[new]
    return pol, soa, state, start
