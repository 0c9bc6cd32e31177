// The reference's ApplyState specs through ClusterUpgradeStateManagerImpl::ApplyState — i.e. through the C ABI and
// the B200 kernel. Needs a GPU; run by tests/test_host_mirror.py::test_reference_specs_on_gpu.
#include "incremental_spec.hpp"

int main() {
  mocks::Runner R;
  bool device_ok = true;
  spec::MakeFn make = [&](upgrade::StateOptions o) {
    std::unique_ptr<upgrade::ClusterUpgradeStateManagerImpl> m;
    if (auto e = upgrade::ClusterUpgradeStateManagerImpl::New(0, o, &m)) {
      std::printf("cannot create manager: %s\n", e->c_str());
      device_ok = false;
      return upgrade::ClusterUpgradeStateManagerImpl::NewDetached(o);
    }
    return m;
  };
  spec::ApplyFn apply = [](spec::Env& e, upgrade::ClusterUpgradeState* s, const upgrade::DriverUpgradePolicySpec* p) {
    return e.m->ApplyState(s, p);
  };
  spec::run(R, make, apply);

  // BuildState (upgrade_state_test.go:115-186) through the manager
  using namespace upgrade;
  auto buildEnv = [&](spec::Env& e, mocks::K8sClientMock& k8s) { e.m = make({}); e.wire(); e.m->K8sClient = &k8s; };
  R.it("BuildState should not fail when no pods exist [upgrade_state_test.go:122-126]", [&] {
    spec::Env e; mocks::K8sClientMock k8s; buildEnv(e, k8s);
    std::unique_ptr<ClusterUpgradeState> st;
    EXPECT(R, !e.m->BuildState("ns", {{"foo", "bar"}}, &st).has_value());
    EXPECT(R, st && st->NodeStates.empty());
  });
  R.it("BuildState should process running daemonset pods, skip unscheduled ones, keep orphans [:128-185]", [&] {
    spec::Env e; mocks::K8sClientMock k8s; buildEnv(e, k8s);
    DaemonSet ds; ds.Name = "ds"; ds.UID = "uid-1"; ds.DesiredNumberScheduled = 2;
    Node n1; n1.Name = "node1"; Node n2; n2.Name = "node2"; n2.Labels[GetUpgradeStateLabelKey()] = UpgradeStateDone;
    e.provider.nodes = {{"node1", &n1}, {"node2", &n2}};
    Pod a; a.Name = "a"; a.NodeName = "node1"; a.Phase = "Running"; a.OwnerReferences = {{"DaemonSet", "ds", "uid-1"}};
    Pod b; b.Name = "b"; b.NodeName = ""; b.Phase = "Pending"; b.OwnerReferences = {{"DaemonSet", "ds", "uid-1"}};
    Pod c; c.Name = "c"; c.NodeName = "node2"; c.Phase = "Running";  // orphan
    // a pod with the driver labels but owned by something that is not a driver DaemonSet: neither GetPodsOwnedbyDs nor
    // GetOrphanedPods keeps it (common_manager.go:190-222)
    Pod d; d.Name = "d"; d.NodeName = "node1"; d.Phase = "Running"; d.OwnerReferences = {{"ReplicaSet", "rs", "3f0e7c1a-9a57-4c55-8f2b-0d6c3f5b7e11"}};
    k8s.daemonSets = {&ds};
    k8s.pods = {&a, &b, &c, &d};
    std::unique_ptr<ClusterUpgradeState> st;
    EXPECT(R, !e.m->BuildState("ns", {}, &st).has_value());
    EXPECT(R, st && st->NodeStates[""].size() == 1 && st->NodeStates[UpgradeStateDone].size() == 1);
    EXPECT(R, st && st->NodeStates[UpgradeStateDone][0]->IsOrphanedPod() && !st->NodeStates[""][0]->IsOrphanedPod());
    ds.DesiredNumberScheduled = 3;  // upgrade_state.go:128-131
    auto err = e.m->BuildState("ns", {}, &st);
    EXPECT(R, err.has_value() && *err == "driver DaemonSet should not have Unscheduled pods");
  });
  // the resourceVersion-keyed encode cache over ust_apply_state_delta_sparse, against ApplyState from scratch
  spec::WorldApplyFn wfull = [](spec::World& w, const upgrade::DriverUpgradePolicySpec* p) { return w.m->ApplyState(&w.state, p); };
  spec::WorldApplyFn wincr = [](spec::World& w, const upgrade::DriverUpgradePolicySpec* p) { return w.m->ApplyStateIncremental(&w.state, p); };
  spec::run_incremental(R, make, wfull, make, wincr, 3000);
  std::printf("# %d passed, %d failed\n", R.passed, R.failed);
  return (R.failed == 0 && device_ok) ? 0 : 1;
}
