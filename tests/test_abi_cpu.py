"""CPU-side checks of libust.so: it loads, exports every entry point include/ust.h declares, refuses to
compute without a device, and its per-policy transition table agrees with the oracle entry by entry."""
import ctypes as C
import re

import numpy as np
import pytest

import helpers
from helpers import abi
from ust import lib as ustlib


def _declared_functions():
    txt = open(abi.HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ust_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ustlib.load()
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libust.so does not export {name}"
    assert sorted(ustlib.EXPORTS) == declared
    # every entry point that takes arguments has its ctypes signature declared (a missing one passes pointers as 32-bit ints)
    for name in declared:
        if name not in ("ust_abi_version", "ust_create_error"):
            assert getattr(lib, name).argtypes is not None, f"{name}: argtypes not declared in ust/lib.py"
    assert lib.ust_abi_version() == abi.UST_ABI_VERSION


def test_struct_layouts_match_header():
    # sizes the C compiler gives include/ust.h (checked against the ctypes mirrors)
    import subprocess, tempfile, os
    src = '#include <stdio.h>\n#include "ust.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(ust_policy), sizeof(ust_counters), sizeof(ust_pods));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.dirname(abi.HEADER), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(abi.Policy), C.sizeof(abi.Counters), C.sizeof(abi.Pods)]


def test_no_cpu_fallback():
    """Without a CUDA device ust_create must fail loudly; nothing computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present; covered by the gpu tests")
    with pytest.raises(ustlib.UstError) as e:
        ustlib.Handle(0)
    assert e.value.code == abi.UST_ERR_CUDA
    assert ustlib.load().ust_apply_state(None, None, 0, None, None, None, None, 0, None, None, None, None, None, None) \
        == abi.UST_ERR_INVALID_ARGUMENT


DERIVED = {"SKIP": 1 << 2, "UNSCHED": 1 << 3, "GRANTED": 1 << 4, "SYNCED": 1 << 9,
           "PD_HAS": 1 << 22, "PD_MISMATCH": 1 << 23, "DRAIN_ERROR": 1 << 24}
KNOWN = abi.UST_F_INPUT_MASK | sum(DERIVED.values())


def _node_for(s, w, pol):
    """SoA fields + workload pods that make the kernel's predicate word equal w for a node in state s."""
    hot = s
    if w & DERIVED["SKIP"]:
        hot |= abi.UST_HOT_SKIP
    if w & DERIVED["UNSCHED"]:
        hot |= abi.UST_HOT_UNSCHEDULABLE
    flags = w & abi.UST_F_INPUT_MASK
    rev = 1 if (w & DERIVED["SYNCED"]) else 2
    pods = []
    rs = abi.UST_PHASE_RUNNING | abi.UST_POD_HAS_CONTROLLER  # replicated, running, no emptyDir: always deletable
    if w & DERIVED["PD_HAS"]:
        pods.append(rs | abi.UST_POD_MATCH_DELETION_FILTER)
    if w & DERIVED["PD_MISMATCH"]:
        if not (w & DERIVED["PD_HAS"]):
            return None  # a mismatch needs at least one pod to delete
        pods.append(rs | abi.UST_POD_MATCH_DELETION_FILTER | abi.UST_POD_MIRROR)  # mirror pods are always kept
    if w & DERIVED["DRAIN_ERROR"]:
        if not pol.drain_delete_emptydir:
            pods.append(rs | abi.UST_POD_HAS_EMPTYDIR | abi.UST_POD_MATCH_DRAIN_SELECTOR)
        elif not pol.drain_force:
            pods.append(abi.UST_PHASE_RUNNING | abi.UST_POD_MATCH_DRAIN_SELECTOR)  # unreplicated
        else:
            return None  # force + deleteEmptyDir: the chain cannot raise an error
    if w & abi.UST_F_WAIT_PODS_RUNNING:
        pods.append(abi.UST_PHASE_PENDING | abi.UST_POD_HAS_CONTROLLER | abi.UST_POD_MATCH_WAIT_SELECTOR)
    else:
        pods.append(abi.UST_PHASE_SUCCEEDED | abi.UST_POD_HAS_CONTROLLER | abi.UST_POD_MATCH_WAIT_SELECTOR)
    return hot, flags, rev, pods


@pytest.mark.parametrize("seed", range(24))
def test_transition_table_matches_oracle(seed):
    """Every reachable (state, window) entry of the product's table == what the reference-shaped oracle
    does to a node built to have exactly those predicates."""
    rng = np.random.default_rng(seed)
    pol = helpers.random_policy(rng)
    pol.evaluate_actuators = 1 if seed % 4 else 0
    lib = ustlib.load()
    for granted in (0, 1):
        # slot budget: unlimited (everything granted) or zero (nothing granted); no abort paths here
        pol.max_parallel_upgrades = 0
        pol.max_unavailable_kind = abi.UST_MAXUNAVAIL_NIL if granted else abi.UST_MAXUNAVAIL_INT
        pol.max_unavailable_value = 0
        if pol.pod_deletion_enabled:
            pol.pod_deletion_spec_present = 1
        hot, flags, rev, off, pf, expect = [], [], [], [0], [], []
        for s in range(16):
            sh = lib.ust_table_window_shift(s)
            for key in range(512):
                w = (key << sh) & 0xFFFFFFFF
                if w & ~KNOWN:
                    continue
                if bool(w & DERIVED["GRANTED"]) != bool(granted):
                    continue
                nd = _node_for(s, w, pol)
                if nd is None:
                    continue
                hot.append(nd[0]); flags.append(nd[1]); rev.append(nd[2]); pf.extend(nd[3]); off.append(len(pf))
                expect.append(lib.ust_table_entry(C.byref(pol), s, w))
        n = len(hot)
        assert n > 800
        # healthy, up-to-date padding so that the unavailable count stays below MaxUnavailable (= total)
        for _ in range(4 * n):
            hot.append(abi.UST_STATE_DONE); flags.append(abi.UST_F_POD_READY); rev.append(1); off.append(len(pf))
        soa = {"state": np.array(hot, np.uint8), "flags": np.array(flags, np.uint32), "pod_rev": np.array(rev, np.int32),
               "ds_idx": np.zeros(len(hot), np.int32), "ds_rev": np.array([1], np.int32)}
        pods = {"pod_off": np.array(off, np.int32), "pod_flags": np.array(pf, np.uint16)}
        rc, nxt, act, oc, cnt = helpers.oracle_apply(pol, soa, pods, variant=0)
        assert rc == 0
        if not pol.use_maintenance_operator:  # requestor mode allocates no slots (upgrade_requestor.go:277-319)
            assert (cnt["upgrades_available"] >= cnt["candidates"]) if granted else (cnt["upgrades_available"] == 0)
        nxt, act, oc = nxt[:n], act[:n], oc[:n]
        exp = np.array(expect, np.uint32)
        assert np.array_equal(nxt, ((exp >> 16) & 0xFF).astype(np.uint8)), np.nonzero(nxt != ((exp >> 16) & 0xFF))[0][:5]
        assert np.array_equal(act, (exp & 0xFFFF).astype(np.uint16)), np.nonzero(act != (exp & 0xFFFF))[0][:5]
        assert np.array_equal(oc, (exp >> 24).astype(np.uint8)), np.nonzero(oc != (exp >> 24))[0][:5]


def test_disabled_policy_table_is_identity():
    lib = ustlib.load()
    pol = abi.make_policy(auto_upgrade=False, max_parallel_upgrades=3)
    for s in range(16):
        for w in (0, 0xFFFFFFFF, 0x12345678):
            for p in (None, C.byref(pol)):
                assert lib.ust_table_entry(p, s, w) == (s << 16) | 0xFF000000


@pytest.mark.parametrize("bits", range(16))
def test_pod_table_256_equals_full_table(bits):
    """The pod-summary kernel looks a pod up by its own eight bits in a 256-byte table and lets the node's selector bit
    decide whether the pod counts (ust_lut.h: ust_build_pod_lut256); gated by the selector-match bits it must say what
    the 2048-entry table (the restatement of the kubectl drain filter chain, pod_manager.go / drain_manager.go) says,
    for every pod_flags value and every policy."""
    fn = ustlib.load().ust_debug_podlut_mismatches
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p]
    pol = abi.make_policy(pod_deletion_enabled=True,
                          pod_deletion={"force": bool(bits & 1), "deleteEmptyDir": bool(bits & 2)},
                          drain={"enable": True, "force": bool(bits & 4), "deleteEmptyDir": bool(bits & 8)},
                          evaluate_actuators=True)
    assert fn(C.addressof(pol)) == 0
