"""Rollout simulation at C3 size: how long `steps` reconciles take on the device, and how the rollout progresses."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "k8s-operator-libs_b200"))
import numpy as np
from ust import abi, lib as ustlib, synth

n = int(os.environ.get("NODES", "10000000"))
steps = int(os.environ.get("STEPS", "200"))
soa = synth.make_nodes(n, synth.CONFIGS["C3"]["seed"])
h = ustlib.Handle(0)
for name, pol in (("C3 policy (maxParallel 100, maxUnavailable 25%)", synth.config_policy("C3")),
                  ("maxParallel 0, maxUnavailable 10%", abi.make_policy(max_parallel_upgrades=0, max_unavailable="10%"))):
    assert h.apply_state(pol, soa, want_outcome=False)[0] == 0
    h.simulate_rollout(pol, n, 3, want_final=False)            # warm-up (tables, hint)
    assert h.apply_state(pol, soa, want_outcome=False)[0] == 0  # back to the original snapshot
    t = time.time()
    rc, done, hist, _ = h.simulate_rollout(pol, n, steps, want_final=False)
    dt = time.time() - t
    print(f"{name}: {steps} reconciles of {n} nodes in {dt * 1e3:.1f} ms = {dt / steps * 1e6:.1f} us per reconcile "
          f"({n * steps / dt:.3g} node-reconciles/s); rc={rc}", flush=True)
    for k in (0, 1, 2, 5, 10, 20, 50, 100, steps - 1):
        if k < steps:
            c = hist[k]
            print(f"  reconcile {k:4d}: upgrade-required {c['hist'][1]:9d} in progress {c['in_progress']:9d} done {c['hist'][11]:9d} "
                  f"failed {c['hist'][12]:7d} slots {c['upgrades_available']:9d}")
