"""Two independently written restatements (reference-shaped objects vs. SoA scalar loop) must agree
bit-for-bit on random snapshots, including abort paths, requestor mode and the pod-list actuators."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import helpers
from helpers import abi


@pytest.mark.parametrize("seed", range(40))
def test_variants_agree_random(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(0, 400))
    p_err = 0.0 if seed % 3 else 0.01
    soa, pods = helpers.random_soa(rng, n, p_err=p_err, with_pods=bool(seed % 2))
    pol = helpers.random_policy(rng)
    a = helpers.oracle_apply(pol, soa, pods, variant=0)
    b = helpers.oracle_apply(pol, soa, pods, variant=1)
    helpers.assert_same(a, b, f"seed {seed}")


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(0, 64), max_par=st.integers(0, 12),
       unav=st.one_of(st.none(), st.integers(0, 70), st.integers(0, 100).map(lambda x: f"{x}%")))
def test_variants_agree_constraint_space(seed, n, max_par, unav):
    """Small clusters, dense sweep of MaxParallelUpgrades x MaxUnavailable (upgrade_inplace.go:49-109)."""
    rng = np.random.default_rng(seed)
    soa, _ = helpers.random_soa(rng, n, all_states=False, wild=False)
    # bias towards upgrade-required / cordon-required so that the slot arithmetic matters
    pick = rng.random(n)
    soa["state"] = np.where(pick < 0.5, (soa["state"] & 0xF0) | 1, np.where(pick < 0.6, (soa["state"] & 0xF0) | 2, soa["state"])).astype(np.uint8)
    pol = helpers.abi.make_policy(max_parallel_upgrades=max_par, max_unavailable=unav)
    a = helpers.oracle_apply(pol, soa, variant=0)
    b = helpers.oracle_apply(pol, soa, variant=1)
    helpers.assert_same(a, b, "constraint sweep")
    # size-independent properties of the slot allocation
    rc, nxt, act, out, cnt = a
    granted = int(np.sum(((soa["state"] & 15) == 1) & (nxt == 2) & ((soa["state"] & helpers.abi.UST_HOT_UNSCHEDULABLE) == 0)))
    assert granted <= max(cnt["upgrades_available"], 0)
    assert cnt["total_managed"] == sum(cnt["hist"][c] for c in (0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12))


def test_disabled_policy_is_noop():
    rng = np.random.default_rng(7)
    soa, _ = helpers.random_soa(rng, 200)
    for variant in (0, 1):
        pol = helpers.abi.make_policy(auto_upgrade=False, max_parallel_upgrades=5)
        rc, nxt, act, out, cnt = helpers.oracle_apply(pol, soa, variant=variant)
        assert rc == 0 and not act.any() and (nxt == (soa["state"] & 15)).all()
        rc2, nxt2, act2, _, _ = helpers.oracle_apply(pol, soa, variant=variant, nil_policy=True)
        assert rc2 == 0 and not act2.any() and (nxt2 == nxt).all()


# ---- rollout simulation: the CPU side ---------------------------------------------------

def test_simulated_rollout_reference_shaped_equals_soa():
    """Both oracle variants drive the same feedback: identical histories and final snapshots; step 0 is a plain
    ApplyState; zero steps change nothing."""
    from ust import synth
    rng = np.random.default_rng(21)
    for n, steps in ((0, 3), (1, 3), (700, 12), (3000, 6)):
        soa, _ = helpers.random_soa(rng, n, wild=False, all_states=True)
        soa["flags"] &= ~np.uint32(abi.UST_F_REQUESTOR_MODE)
        pol = abi.make_policy(max_parallel_upgrades=int(rng.integers(0, 50)), max_unavailable=["25%", None, 40][int(rng.integers(0, 3))],
                              pod_deletion_enabled=bool(rng.integers(0, 2)), validation_enabled=bool(rng.integers(0, 2)),
                              pod_deletion={"force": False}, drain={"enable": True},
                              wait_for_completion={"podSelector": "app=job", "timeoutSeconds": 30})
        a = helpers.oracle_simulate(pol, soa, steps, variant=0)
        b = helpers.oracle_simulate(pol, soa, steps, variant=1)
        assert a[0] == b[0] == 0 and a[1] == b[1] == steps and a[2] == b[2]
        for k in ("state", "flags", "pod_rev"):
            assert np.array_equal(a[3][k], b[3][k]), k
        if steps:
            one = helpers.oracle_apply(_with_actuators(pol), soa, variant=1)
            assert a[2][0] == one[4]
        z = helpers.oracle_simulate(pol, soa, 0, variant=1)
        assert z[1] == 0 and all(np.array_equal(z[3][k], soa[k]) for k in ("state", "flags", "pod_rev"))


def _with_actuators(pol):
    import copy
    p = copy.copy(pol)
    p.evaluate_actuators = 1
    return p


def test_simulated_rollout_converges_and_respects_the_budget():
    """With ideal actuators the budget is the only brake: every reconcile moves at most max(upgradesAvailable, 0)
    nodes out of upgrade-required (plus the already cordoned ones, upgrade_inplace.go:87-101), the unavailable count
    never exceeds what it started with or maxUnavailable allows, and the rollout ends with every node done, failed or
    deliberately skipped."""
    from ust import synth
    soa = synth.make_nodes(20_000, 5)
    pol = abi.make_policy(max_parallel_upgrades=0, max_unavailable="10%")
    rc, done, hist, fin = helpers.oracle_simulate(pol, soa, 120)
    assert rc == 0 and done == 120
    code = fin["state"] & 15
    skip = (fin["state"] & abi.UST_HOT_SKIP) != 0
    left = np.isin(code, [1]) & ~skip
    assert not left.any(), "an unskipped node is still waiting for a slot after 120 reconciles"
    assert set(np.unique(code)) <= {1, 11, 12, 13, 14, 15}
    for k in range(1, 120):
        moved = hist[k - 1]["hist"][1] - hist[k]["hist"][1]
        assert moved <= max(hist[k - 1]["upgrades_available"], 0) + hist[k - 1]["unavailable"], k
        if hist[k]["max_unavailable"] < hist[k]["total_managed"] and k > 8:
            assert hist[k]["unavailable"] <= max(hist[0]["unavailable"], hist[k]["max_unavailable"]), k


def test_oracle_timed_simulation_follows_the_wait_start_vectors():
    """The oracle's restatement of the timed rollout simulation against the reference's wait-for-completion timeout rule
    (pod_manager.go:331-368, vectors pod_manager_test.go:183-229), reconcile by reconcile (helpers.wait_timeout_timeline):
    annotation set on the first reconcile, nothing until now > start + timeout, then pod-deletion-required."""
    G = helpers.load_golden()
    for timeout, dt, steps in ((100, 30, 8), (45, 45, 5), (10, 60, 3)):
        pol, soa, state, start = helpers.wait_timeout_timeline(G["daemonset_hash"], timeout, dt, steps)
        opt = abi.SimOptions(dt, timeout, 10 ** 6, 0, 600, 0)
        for variant in (0, 1):
            rc, done, hist, fin = helpers.oracle_simulate_timed(pol, opt, soa, steps, variant=variant)
            assert rc == 0 and done == steps
            assert [abi.STATE_NAMES[c & 15] for c in fin["state"]] == state, (timeout, dt, state)
            assert all(((fin["flags"][i] & abi.UST_F_WAIT_START_ANNO) != 0) == (start[i] is not None) for i in range(3))
        # first reconcile with now > start + timeout
        k_out = next(k for k in range(steps + 50) if k * dt > timeout)
        if k_out < steps:
            assert hist[k_out]["hist"][abi.UST_STATE_WAIT_FOR_JOBS_REQUIRED] == 3
            if k_out + 1 < steps:
                assert hist[k_out + 1]["hist"][abi.UST_STATE_WAIT_FOR_JOBS_REQUIRED] == 0
