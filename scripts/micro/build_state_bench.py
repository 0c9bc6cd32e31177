"""Device time of the BuildState owner join (ust_build_state_uids) at 10 M driver pods, kernels only (ncu launch list)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "k8s-operator-libs_b200"))
import numpy as np
from ust import lib as ustlib

n, n_ds = int(os.environ.get("PODS", "10000000")), 4
rng = np.random.default_rng(1)
ds_uid = rng.integers(1, 2 ** 63, size=(n_ds, 2), dtype=np.uint64)
truth = rng.integers(0, n_ds, n).astype(np.int32)
owner = ds_uid[truth]
state = rng.integers(0, 13, n).astype(np.uint8)
desired = np.bincount(truth, minlength=n_ds).astype(np.int32)
h = ustlib.Handle(0)
for _ in range(3):
    t = time.time(); rc, ds_idx, cnt = h.build_state_uids(state, owner, ds_uid, desired); dt = time.time() - t
    assert rc == 0 and np.array_equal(ds_idx, truth)
    print("ust_build_state_uids end to end (pageable host arrays): %.1f ms" % (dt * 1e3), flush=True)
