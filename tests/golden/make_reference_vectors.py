#!/usr/bin/env python
"""Hand transcription of the reference's own known-answer tests for the ApplyState / BuildState path
into reference_vectors.json (run this script to regenerate the JSON; it needs nothing but Python).

Every vector cites the Go test it restates (`ref`, path relative to the reference repo @ 11e747a).
`expect_counts` / per-node `expect` are EXACTLY the assertions the Go test makes — nothing derived.
Where the Go test only counts states, `derived` additionally carries the per-node answer worked out
by hand from upgrade_inplace.go:71-109 (slice order); tests treat `derived` as a weaker, clearly
labelled check.

Node fields (all optional):
  state            value of the upgrade-state label ("" = unknown)
  unschedulable    Spec.Unschedulable
  ready            status of the NodeReady condition ("True"/"False"/"Unknown"); absent = no condition
  skip             value of the skip label
  anno             {upgrade-requested|safe-load|initial-state|requestor-mode|wait-start: value}
  pod              null (NodeUpgradeState.DriverPod nil) or
                   {hash, phase, containers:[[ready,restarts]..], init:[[ready,restarts]..], terminating}
                   hash absent  => no controller-revision-hash label. The suite's PodManager mock returns
                   "" for a missing label and never errors (upgrade_suit_test.go:159-168): vectors that
                   rely on that say mock_hash_getter=true and are encoded with hash "".
  ds               true when DriverDaemonSet != nil (its revision hash is the mock's constant
                   "test-hash-12345", upgrade_suit_test.go:169-171)
  nm               null or {"ready": bool}   (NodeMaintenance object / Ready condition)
  validation_done  what ValidationManager.Validate returns for the node (mock: true, :179-182)
  workload         list of workload pods for actuator vectors:
                   {phase, controller:null|"ReplicaSet"|"DaemonSet", ds_missing, mirror, emptydir,
                    match_filter, match_wait, match_drain}
"""
import json
import os

H = "test-hash-12345"
UP = {"hash": H}                                  # upToDatePod without status
UP_RUN = {"hash": H, "phase": "Running"}
OUT = {"hash": "test-hash-outdated"}
OUT_RUN = {"hash": "test-hash-outdated", "phase": "Running"}
READY = {"hash": H, "phase": "Running", "containers": [[True, 0]]}
AUTO = {"autoUpgrade": True}


def n(state, **kw):
    d = {"state": state}
    d.update(kw)
    return d


def ur(**kw):
    return n("upgrade-required", **kw)


V = []


def vec(name, ref, policy, nodes, **kw):
    d = {"name": name, "ref": ref, "policy": policy, "nodes": nodes}
    d.update(kw)
    V.append(d)


F = "pkg/upgrade/upgrade_state_test.go"

vec("nil currentState fails", F + ":190-192", {}, [], nil_state=True, expect_error="NIL_STATE")
vec("nil upgradePolicy succeeds as no-op", F + ":193-195", None, [], expect_error=None)

vec("up-to-date -> done, outdated -> upgrade-required", F + ":196-225", AUTO, [
    n("", pod=UP, ds=True, expect={"state": "upgrade-done"}),
    n("", pod=OUT, ds=True, expect={"state": "upgrade-required"}),
    n("upgrade-done", pod=UP, ds=True, expect={"state": "upgrade-done"}),
    n("upgrade-done", pod=OUT, ds=True, expect={"state": "upgrade-required"}),
])

vec("outdated unschedulable nodes get the initial-state annotation", F + ":226-270", AUTO, [
    n("", pod=UP, ds=True, expect={"state": "upgrade-done", "anno_absent": ["initial-state"]}),
    n("", pod=OUT, ds=True, unschedulable=True, expect={"state": "upgrade-required", "anno_present": ["initial-state"]}),
    n("upgrade-done", pod=UP, ds=True, expect={"state": "upgrade-done", "anno_absent": ["initial-state"]}),
    n("upgrade-done", pod=OUT, ds=True, unschedulable=True,
      expect={"state": "upgrade-required", "anno_present": ["initial-state"]}),
])

vec("safe-load annotation on an up-to-date done node -> upgrade-required", F + ":271-293", AUTO, [
    n("upgrade-done", pod=UP, ds=True, anno={"safe-load": "true"}, expect={"state": "upgrade-required"}),
])

vec("maxParallel 0 schedules all", F + ":294-320", {"autoUpgrade": True, "maxParallelUpgrades": 0},
    [ur() for _ in range(5)],
    expect_counts={"upgrade-required": 0, "cordon-required": 5})

vec("maxParallel 3 of 5", F + ":321-349", {"autoUpgrade": True, "maxParallelUpgrades": 3},
    [ur() for _ in range(5)],
    expect_counts={"upgrade-required": 2, "cordon-required": 3},
    derived=["cordon-required"] * 3 + ["upgrade-required"] * 2)

vec("maxParallel 4 with 3 already cordon-required", F + ":350-383",
    {"autoUpgrade": True, "maxParallelUpgrades": 4, "drain": {"enable": True}},
    [ur(), ur()] + [n("cordon-required") for _ in range(3)],
    expect_counts={"upgrade-required": 1},
    expect_count_sums=[[["cordon-required", "wait-for-jobs-required"], 4]],
    derived=["cordon-required", "upgrade-required"] + ["wait-for-jobs-required"] * 3)

vec("maxParallel 0, maxUnavailable 100%", F + ":384-412",
    {"autoUpgrade": True, "maxParallelUpgrades": 0, "maxUnavailable": "100%"},
    [ur(), ur(), ur(), ur(unschedulable=True), ur(unschedulable=True)],
    expect_counts={"upgrade-required": 0, "cordon-required": 5})

vec("maxParallel 0, maxUnavailable 50%, two already cordoned", F + ":413-440",
    {"autoUpgrade": True, "maxParallelUpgrades": 0, "maxUnavailable": "50%"},
    [ur(), ur(), ur(), ur(unschedulable=True), ur(unschedulable=True)],
    expect_counts={"upgrade-required": 2, "cordon-required": 3},
    derived=["cordon-required", "upgrade-required", "upgrade-required", "cordon-required", "cordon-required"])

vec("maxUnavailable 50% with two unavailable done nodes", F + ":441-513",
    {"autoUpgrade": True, "maxParallelUpgrades": 0, "maxUnavailable": "50%"},
    [ur(), ur(), ur(),
     n("upgrade-done", unschedulable=True, pod=UP_RUN, ds=True),
     n("upgrade-done", unschedulable=True, pod=UP_RUN, ds=True)],
    expect_counts={"upgrade-done": 2, "cordon-required": 1, "upgrade-required": 2},
    expect_restart=[],
    derived=["cordon-required", "upgrade-required", "upgrade-required", "upgrade-done", "upgrade-done"])

vec("maxParallel 3 and maxUnavailable 2", F + ":514-544",
    {"autoUpgrade": True, "maxParallelUpgrades": 3, "maxUnavailable": 2},
    [ur() for _ in range(5)],
    expect_counts={"upgrade-required": 3, "cordon-required": 2},
    derived=["cordon-required"] * 2 + ["upgrade-required"] * 3)

for ref in (":545-579", ":580-614"):
    vec("maxParallel 4 and maxUnavailable 4 with 3 cordon-required", F + ref,
        {"autoUpgrade": True, "maxParallelUpgrades": 4, "maxUnavailable": 4, "drain": {"enable": True}},
        [ur(), ur()] + [n("cordon-required") for _ in range(3)],
        expect_counts={"upgrade-required": 1},
        expect_count_sums=[[["cordon-required", "wait-for-jobs-required"], 4]],
        derived=["cordon-required", "upgrade-required"] + ["wait-for-jobs-required"] * 3)

vec("wait-for-jobs without pod deletion filter -> drain-required", F + ":615-633", AUTO,
    [n("wait-for-jobs-required", expect={"state": "drain-required"}) for _ in range(3)])

vec("wait-for-jobs with pod deletion filter -> pod-deletion-required", F + ":634-657", AUTO,
    [n("wait-for-jobs-required", expect={"state": "pod-deletion-required"}) for _ in range(3)],
    options={"podDeletionEnabled": True})

vec("pod-deletion-required with pod deletion disabled -> drain-required, no eviction", F + ":658-695", AUTO,
    [n("pod-deletion-required", expect={"state": "drain-required"}) for _ in range(3)],
    expect_eviction=[])

vec("drain-required with nil drain spec -> pod-restart-required", F + ":696-718", AUTO,
    [n("drain-required", expect={"state": "pod-restart-required"}) for _ in range(3)])
vec("drain-required with disabled drain spec -> pod-restart-required", F + ":720-729",
    {"autoUpgrade": True, "drain": {"enable": False}},
    [n("drain-required", expect={"state": "pod-restart-required"}) for _ in range(3)])

vec("drain enabled schedules drain for all three nodes", F + ":730-763",
    {"autoUpgrade": True, "drain": {"enable": True}},
    [n("drain-required", expect={"state": "drain-required"}) for _ in range(3)],
    expect_drain=[0, 1, 2])

vec("drain manager error fails ApplyState", F + ":764-788",
    {"autoUpgrade": True, "drain": {"enable": True}},
    [n("drain-required") for _ in range(3)],
    actuator_error="ScheduleNodesDrain", expect_error="ACTUATOR")

vec("restart only the outdated, non-terminating pod", F + ":789-849", AUTO, [
    n("pod-restart-required", pod=UP_RUN, ds=True),
    n("pod-restart-required", pod=OUT_RUN, ds=True),
    n("pod-restart-required", pod={"hash": "test-hash-outdated", "terminating": True}, ds=True),
], expect_restart=[1])

vec("safe-load annotation is removed instead of restarting", F + ":850-883", AUTO, [
    n("pod-restart-required", pod=UP_RUN, ds=True, anno={"safe-load": "true"},
      expect={"anno_absent": ["safe-load"]}),
], expect_restart=[])

vec("ready + up-to-date in pod-restart / upgrade-failed -> uncordon-required", F + ":884-919", AUTO, [
    n("pod-restart-required", pod=READY, ds=True, expect={"state": "uncordon-required"}),
    n("upgrade-failed", pod=READY, ds=True, expect={"state": "uncordon-required"}),
])

vec("... and initially unschedulable -> upgrade-done, annotation removed", F + ":920-971", AUTO, [
    n("pod-restart-required", pod=READY, ds=True, unschedulable=True, anno={"initial-state": "true"},
      expect={"state": "upgrade-done", "anno_absent": ["initial-state"]}),
    n("upgrade-failed", pod=READY, ds=True, unschedulable=True, anno={"initial-state": "true"},
      expect={"state": "upgrade-done", "anno_absent": ["initial-state"]}),
])

vec("repeated container restarts -> upgrade-failed", F + ":972-1017", AUTO, [
    n("pod-restart-required", ds=True, pod={"hash": H, "phase": "Running", "containers": [[False, 0]]},
      expect={"state": "pod-restart-required"}),
    n("pod-restart-required", ds=True,
      pod={"hash": H, "phase": "Running", "containers": [[False, 0]], "init": [[True, 0]]},
      expect={"state": "pod-restart-required"}),
    n("pod-restart-required", ds=True,
      pod={"hash": H, "phase": "Running", "containers": [[False, 11]], "init": [[True, 0]]},
      expect={"state": "upgrade-failed"}),
    n("pod-restart-required", ds=True,
      pod={"hash": H, "phase": "Running", "containers": [[False, 0]], "init": [[False, 11]]},
      expect={"state": "upgrade-failed"}),
])

vec("validation enabled: ready pod -> validation-required", F + ":1018-1053", AUTO, [
    n("pod-restart-required", pod=READY, ds=True, expect={"state": "validation-required"}),
], options={"validationEnabled": True})

vec("validation done -> uncordon-required", F + ":1054-1088", AUTO, [
    n("validation-required", pod={"hash": ""}, ds=True, validation_done=True, expect={"state": "uncordon-required"}),
], options={"validationEnabled": True}, mock_hash_getter=True)

vec("validation done, initially unschedulable -> upgrade-done", F + ":1089-1127", AUTO, [
    n("validation-required", pod={"hash": ""}, ds=True, validation_done=True, unschedulable=True,
      anno={"initial-state": "true"}, expect={"state": "upgrade-done", "anno_absent": ["initial-state"]}),
], options={"validationEnabled": True}, mock_hash_getter=True)

vec("uncordon-required -> upgrade-done", F + ":1128-1153", AUTO, [
    n("uncordon-required", expect={"state": "upgrade-done"}),
], expect_uncordon=[0])

vec("cordon manager failure fails ApplyState", F + ":1154-1178", AUTO, [
    n("uncordon-required", expect={"state_not": "upgrade-done"}),
], actuator_error="Uncordon", expect_error="ACTUATOR")

vec("orphaned pod: unknown/done stay done", F + ":1180-1199", AUTO, [
    n("", pod={}, ds=False, expect={"state": "upgrade-done"}),
    n("upgrade-done", pod={}, ds=False, expect={"state": "upgrade-done"}),
])

vec("orphaned pod + upgrade-requested -> upgrade-required", F + ":1200-1221", AUTO, [
    n("", pod={}, ds=False, anno={"upgrade-requested": "true"}, expect={"state": "upgrade-required"}),
    n("upgrade-done", pod={}, ds=False, anno={"upgrade-requested": "true"}, expect={"state": "upgrade-required"}),
])

vec("upgrade-required orphan: cordon-required and upgrade-requested removed", F + ":1222-1237", AUTO, [
    ur(pod={}, ds=False, anno={"upgrade-requested": "true"},
       expect={"state": "cordon-required", "anno_absent": ["upgrade-requested"]}),
])

vec("orphaned pod in pod-restart-required is restarted", F + ":1238-1267", AUTO, [
    n("pod-restart-required", pod=OUT_RUN, ds=False),
], expect_restart=[0])

vec("upgrade-failed with a pod lacking the hash label stays failed (mock hash getter)", F + ":1268-1294", AUTO, [
    n("upgrade-failed", ds=True, pod={"hash": "", "phase": "Running", "containers": [[True, 0]]},
      expect={"state": "upgrade-failed"}),
], mock_hash_getter=True)

# ---- requestor mode (upgrade_requestor.go), policy-level option useMaintenanceOperator ------------
RQ = {"useMaintenanceOperator": True}
DRAIN_ON = {"autoUpgrade": True, "drain": {"enable": True}}
NOHASH = {"hash": ""}  # NewPod(...) builder sets no controller-revision-hash label; mock getter => ""

vec("requestor: upgrade-required -> node-maintenance-required + requestor-mode annotation", F + ":1296-1352",
    DRAIN_ON,
    [ur(pod=NOHASH, ds=True, expect={"anno_present": ["requestor-mode"]}) for _ in range(3)],
    options=RQ, mock_hash_getter=True,
    derived=["node-maintenance-required"] * 3, expect_nm_change=[0, 1, 2])

vec("requestor (shared): upgrade-required with existing NodeMaintenance", F + ":1354-1392",
    DRAIN_ON,
    [ur(pod=NOHASH, ds=True, nm={"ready": False}, expect={"anno_present": ["requestor-mode"]}) for _ in range(3)],
    options=RQ, mock_hash_getter=True, derived=["node-maintenance-required"] * 3)

vec("requestor: node-maintenance-required stays while NodeMaintenance is not Ready", F + ":1393-1433",
    DRAIN_ON,
    [n("node-maintenance-required", pod=NOHASH, ds=True, nm={"ready": False},
       expect={"state": "node-maintenance-required"})] +
    [n("node-maintenance-required", pod=NOHASH, ds=True, nm={"ready": False}) for _ in range(2)],
    options=RQ, mock_hash_getter=True)

vec("requestor: NodeMaintenance Ready -> pod-restart-required", F + ":1435-1483", DRAIN_ON,
    [n("node-maintenance-required", pod=NOHASH, ds=True, nm={"ready": True}, anno={"upgrade-requested": "true"},
       expect={"state": "pod-restart-required"})],
    options=RQ, mock_hash_getter=True)

vec("requestor: NodeMaintenance missing -> upgrade-required", F + ":1485-1510", DRAIN_ON,
    [n("node-maintenance-required", pod=NOHASH, ds=True, nm=None, anno={"upgrade-requested": "true"},
       expect={"state": "upgrade-required"})],
    options=RQ, mock_hash_getter=True)

vec("requestor: cordon-required keeps the in-place flow -> wait-for-jobs-required", F + ":1512-1530", DRAIN_ON,
    [n("cordon-required", pod=NOHASH, ds=True, nm={"ready": False}, expect={"state": "wait-for-jobs-required"})] +
    [n("cordon-required", pod=NOHASH, ds=True, nm={"ready": False}) for _ in range(2)],
    options=RQ, mock_hash_getter=True)

vec("requestor: uncordon-required with requestor-mode annotation -> upgrade-done", F + ":1532-1564", DRAIN_ON,
    [n("uncordon-required", pod=NOHASH, ds=True, nm={"ready": False},
       anno={"upgrade-requested": "true", "requestor-mode": "true"},
       expect={"state": "upgrade-done", "anno_absent": ["requestor-mode", "initial-state"]}) for _ in range(3)],
    options=RQ, mock_hash_getter=True, expect_nm_change=[0, 1, 2])

vec("requestor (shared): validation-required + initial-state + requestor-mode -> uncordon-required", F + ":1566-1609",
    {"autoUpgrade": True, "drain": {"enable": False}},
    [n("validation-required", pod=NOHASH, ds=True, nm={"ready": False},
       anno={"requestor-mode": "true", "initial-state": "true"},
       expect={"state": "uncordon-required", "anno_present": ["requestor-mode"], "anno_absent": ["initial-state"]})
     for _ in range(3)],
    options=RQ, mock_hash_getter=True)

vec("requestor (not owning NM): validation-required -> uncordon-required", F + ":1611-1647", DRAIN_ON,
    [n("validation-required", pod=NOHASH, ds=True, nm={"ready": False},
       anno={"requestor-mode": "true", "initial-state": "true"}, expect={"state": "uncordon-required"})] +
    [n("validation-required", pod=NOHASH, ds=True, nm={"ready": False},
       anno={"requestor-mode": "true", "initial-state": "true"}) for _ in range(2)],
    options=RQ, mock_hash_getter=True)

vec("requestor (shared): uncordon-required -> upgrade-done, requestor-mode annotation removed", F + ":1649-1696",
    DRAIN_ON,
    [n("uncordon-required", pod=NOHASH, ds=True, nm={"ready": False},
       anno={"requestor-mode": "true", "initial-state": "true"},
       expect={"state": "upgrade-done", "anno_absent": ["requestor-mode"]}) for _ in range(3)],
    options=RQ, mock_hash_getter=True)

vec("requestor: orphaned pod + upgrade-requested -> upgrade-required", F + ":1698-1721", AUTO, [
    n("", pod={}, ds=False, anno={"upgrade-requested": "true"}, expect={"state": "upgrade-required"}),
    n("upgrade-done", pod={}, ds=False, anno={"upgrade-requested": "true"}, expect={"state": "upgrade-required"}),
], options=RQ)

vec("requestor: safe-load annotation -> upgrade-required", F + ":1723-1745", AUTO, [
    n("upgrade-done", pod=UP, ds=True, anno={"safe-load": "true"}, expect={"state": "upgrade-required"}),
], options=RQ)

vec("requestor: validation done, initially unschedulable, no requestor annotation -> upgrade-done",
    F + ":1746-1784", AUTO, [
        n("validation-required", pod={"hash": ""}, ds=True, validation_done=True, unschedulable=True,
          anno={"initial-state": "true"}, expect={"state": "upgrade-done", "anno_absent": ["initial-state"]}),
    ], options={"useMaintenanceOperator": True, "validationEnabled": True}, mock_hash_getter=True)

# ---- actuator decisions pinned by pod_manager_test.go (evaluate_actuators) -------------------------
P = "pkg/upgrade/pod_manager_test.go"
STANDALONE = {"phase": "Running", "controller": None}
CPU = dict(STANDALONE, match_filter=False)
GPU = dict(STANDALONE, match_filter=True)
GPU_ED = dict(GPU, emptydir=True)


def act(name, ref, policy, node, options=None):
    vec(name, ref, policy, [node], options=options or {}, evaluate_actuators=True)


WAIT = {"autoUpgrade": True, "waitForCompletion": {"podSelector": "app=my-app", "timeoutSeconds": 0}}
WAIT30 = {"autoUpgrade": True, "waitForCompletion": {"podSelector": "app=my-app", "timeoutSeconds": 30}}
act("wait-for-jobs: workload Succeeded -> pod-deletion-required, no start-time annotation", P + ":120-153", WAIT,
    n("wait-for-jobs-required", workload=[{"phase": "Succeeded", "controller": None, "match_wait": True}],
      expect={"outcome": "pod-deletion-required", "actions_absent": ["SET_WAIT_START"]}))
act("wait-for-jobs: workload Running, no timeout -> unchanged, no annotation", P + ":154-182", WAIT,
    n("wait-for-jobs-required", workload=[{"phase": "Running", "controller": None, "match_wait": True}],
      expect={"outcome": "wait-for-jobs-required", "actions_absent": ["SET_WAIT_START", "CLEAR_WAIT_START"]}))
act("wait-for-jobs: Running with timeout, first pass adds the start-time annotation", P + ":183-214", WAIT30,
    n("wait-for-jobs-required", workload=[{"phase": "Running", "controller": None, "match_wait": True}],
      expect={"outcome": "wait-for-jobs-required", "actions_present": ["SET_WAIT_START"]}))
act("wait-for-jobs: Running, start-time 35 s ago, timeout 30 -> pod-deletion-required, annotation removed",
    P + ":216-229", WAIT30,
    n("wait-for-jobs-required", anno={"wait-start": "now-35"},
      workload=[{"phase": "Running", "controller": None, "match_wait": True}],
      expect={"outcome": "pod-deletion-required", "actions_present": ["CLEAR_WAIT_START"]}))


def pd(force, ded, drain):
    return {"autoUpgrade": True, "podDeletion": {"force": force, "deleteEmptyDir": ded},
            "drain": {"enable": drain}}


PDE = {"podDeletionEnabled": True}
act("eviction: standalone gpu pods with force -> pod-restart-required", P + ":236-265", pd(True, False, False),
    n("pod-deletion-required", workload=[CPU, GPU, GPU], expect={"outcome": "pod-restart-required"}), PDE)
act("eviction: no force, drain disabled -> upgrade-failed", P + ":267-298", pd(False, False, False),
    n("pod-deletion-required", workload=[CPU, GPU, GPU], expect={"outcome": "upgrade-failed"}), PDE)
act("eviction: no force, drain enabled -> drain-required", P + ":300-330", pd(False, False, True),
    n("pod-deletion-required", workload=[CPU, GPU, GPU], expect={"outcome": "drain-required"}), PDE)
act("eviction: force + deleteEmptyDir with an emptyDir pod -> pod-restart-required", P + ":332-364",
    pd(True, True, False),
    n("pod-deletion-required", workload=[CPU, GPU, GPU, GPU_ED], expect={"outcome": "pod-restart-required"}), PDE)
act("eviction: force, emptyDir pod, deleteEmptyDir=false, drain disabled -> upgrade-failed", P + ":366-397",
    pd(True, False, False),
    n("pod-deletion-required", workload=[CPU, GPU_ED], expect={"outcome": "upgrade-failed"}), PDE)
act("eviction: force, emptyDir pod, deleteEmptyDir=false, drain enabled -> drain-required", P + ":399-429",
    pd(True, False, True),
    n("pod-deletion-required", workload=[CPU, GPU_ED], expect={"outcome": "drain-required"}), PDE)

# ---- BuildState (upgrade_state_test.go:115-186) ----------------------------------------------------
B = []
B.append({"name": "no pods", "ref": F + ":122-126", "daemonsets": [], "pods": [],
          "expect_error": None, "expect_buckets": {}})
B.append({"name": "running daemonset pod", "ref": F + ":128-148", "daemonsets": [{"desired": 1}],
          "pods": [{"ds": 0, "node_state": "", "node_name": "node", "phase": "Running"}],
          "expect_error": None, "expect_buckets": {"": 1}})
B.append({"name": "daemonset pod not scheduled yet is skipped", "ref": F + ":150-172", "daemonsets": [{"desired": 1}],
          "pods": [{"ds": 0, "node_state": "", "node_name": "", "phase": "Pending"}],
          "expect_error": None, "expect_buckets": {}})
B.append({"name": "orphaned pod", "ref": F + ":174-185", "daemonsets": [],
          "pods": [{"ds": None, "node_state": "", "node_name": "node", "phase": "Running"}],
          "expect_error": None, "expect_buckets": {"": 1}, "expect_orphan": [0]})
# upgrade_state.go:128-131 has no Go test; restated from the code, labelled as such
B.append({"name": "daemonset with fewer pods than DesiredNumberScheduled errors (code-derived, unpinned)",
          "ref": "pkg/upgrade/upgrade_state.go:128-131", "daemonsets": [{"desired": 2}],
          "pods": [{"ds": 0, "node_state": "", "node_name": "node", "phase": "Running"}],
          "expect_error": "DS_UNSCHEDULED", "expect_buckets": None})

# ---- intstr scaling pinned through the vectors above ------------------------------------------------
S = [
    {"ref": F + ":384-412", "percent": 100, "total": 5, "expect": 5},
    {"ref": F + ":413-440", "percent": 50, "total": 5, "expect": 3},
    {"ref": F + ":441-513", "percent": 50, "total": 5, "expect": 3},
]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump({"reference": "NVIDIA/k8s-operator-libs @ 11e747a", "daemonset_hash": H,
                   "apply_state": V, "build_state": B, "intstr": S}, f, indent=1)
    print("wrote", out, len(V), "ApplyState vectors,", len(B), "BuildState vectors")
