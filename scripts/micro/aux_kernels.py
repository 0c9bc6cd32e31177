"""One call of each auxiliary kernel at 10 M units, for an ncu launch list (profiles/README.md)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "k8s-operator-libs_b200"))
import numpy as np
from ust import abi, lib as ustlib, synth

n = 10_000_000
soa = synth.make_nodes(n, synth.CONFIGS["C3"]["seed"])
pol = synth.config_policy("C3")
h = ustlib.Handle(0)
assert h.apply_state(pol, soa, want_outcome=False)[0] == 0
rng = np.random.default_rng(1)
idx = rng.choice(n, size=n // 100, replace=False).astype(np.int64)
assert h.apply_state_delta(pol, n, idx, {k: soa[k][idx] for k in ("state", "flags", "pod_rev", "ds_idx")}, soa["ds_rev"], want_outcome=False)[0] == 0
assert h.simulate_rollout(pol, n, 3, want_final=False)[0] == 0
ds_idx = rng.integers(0, 4, n).astype(np.int32)
state = rng.integers(0, 13, n).astype(np.uint8)
rc, _ = h.build_state(state, ds_idx, np.bincount(ds_idx, minlength=4).astype(np.int32))
assert rc == 0
print("ok")
