"""Two independently written restatements (reference-shaped objects vs. SoA scalar loop) must agree
bit-for-bit on random snapshots, including abort paths, requestor mode and the pod-list actuators."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import helpers


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
