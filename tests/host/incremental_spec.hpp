// A multi-reconcile scenario for ClusterUpgradeStateManagerImpl::ApplyStateIncremental (the resourceVersion-keyed encode
// cache): two identical worlds, one reconciled with ApplyState (everything re-encoded and re-evaluated from
// scratch, the reference's way), one with ApplyStateIncremental. After every reconcile the two worlds must be
// indistinguishable: same error, same labels / annotations on every node, same actuator calls. Between reconciles the
// world moves the way a cluster does: cordons take effect, restarted driver pods come back at the current revision,
// operators edit labels - every change bumps the object's resourceVersion, as the API server would.
#pragma once
#include <set>

#include "upgrade_state_spec.hpp"

namespace spec {

struct VersionedProvider : NodeUpgradeStateProviderMock {
  static void bump(Node* n) { n->ResourceVersion = std::to_string(std::stoll(n->ResourceVersion.empty() ? "0" : n->ResourceVersion) + 1); }
  Error ChangeNodeUpgradeState(Node* node, const std::string& s) override { bump(node); return NodeUpgradeStateProviderMock::ChangeNodeUpgradeState(node, s); }
  Error ChangeNodeUpgradeAnnotation(Node* node, const std::string& k, const std::string& v) override {
    bump(node);
    return NodeUpgradeStateProviderMock::ChangeNodeUpgradeAnnotation(node, k, v);
  }
};

struct World {
  std::unique_ptr<ClusterUpgradeStateManagerImpl> m;
  VersionedProvider provider;
  CordonManagerMock cordon;
  DrainManagerMock drain;
  PodManagerMock pods;
  ValidationManagerMock validation;
  SafeDriverLoadManagerImpl safeLoad{&provider};
  std::deque<Node> nodes;
  std::deque<Pod> podObjs;
  std::deque<NodeUpgradeState> entries;
  DaemonSet daemonSet;
  ClusterUpgradeState state;
  void wire() {
    m->NodeUpgradeStateProvider = &provider; m->DrainManager = &drain; m->CordonManager = &cordon;
    m->PodManager = &pods; m->ValidationManager = &validation; m->SafeDriverLoadManager = &safeLoad;
  }
  // BuildState's result for the current labels: buckets by label, node order within a bucket = list order
  void snapshot() {
    state = ClusterUpgradeState();
    entries.clear();
    for (size_t i = 0; i < nodes.size(); i++) {
      if (nodes[i].Name.empty()) continue;  // a node that left the cluster
      entries.emplace_back();
      NodeUpgradeState& e = entries.back();
      e.Node = &nodes[i];
      e.ListIndex = (int64_t)i;  // what BuildState stamps: the position in its pod list
      e.DriverPod = &podObjs[i];
      e.DriverDaemonSet = (i % 17 == 3) ? nullptr : &daemonSet;  // a few orphaned pods
      state.NodeStates[getNodeUpgradeState(&nodes[i])].push_back(&e);
    }
  }
};

struct Lcg {
  uint64_t s;
  uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
  bool chance(int pct) { return next() % 100 < (uint32_t)pct; }
};

inline void populate(World& w, int n, uint64_t seed) {
  const char* states[] = {"", UpgradeStateUpgradeRequired, UpgradeStateCordonRequired, UpgradeStateWaitForJobsRequired,
                          UpgradeStatePodDeletionRequired, UpgradeStateDrainRequired, UpgradeStatePodRestartRequired,
                          UpgradeStateValidationRequired, UpgradeStateUncordonRequired, UpgradeStateDone, UpgradeStateFailed,
                          UpgradeStateDone, UpgradeStateDone, UpgradeStateUpgradeRequired, "some-other-label"};
  Lcg r{seed};
  w.daemonSet.Name = "driver"; w.daemonSet.UID = "ds-uid-1"; w.daemonSet.ResourceVersion = "1";
  for (int i = 0; i < n; i++) {
    w.nodes.emplace_back();
    Node& nd = w.nodes.back();
    nd.Name = "node-" + std::to_string(i);
    nd.ResourceVersion = "1";
    nd.Labels[GetUpgradeStateLabelKey()] = states[r.next() % (sizeof(states) / sizeof(states[0]))];
    nd.Unschedulable = r.chance(15);
    if (r.chance(3)) nd.Labels[GetUpgradeSkipNodeLabelKey()] = "true";
    if (r.chance(3)) nd.Annotations[GetUpgradeRequestedAnnotationKey()] = "true";
    if (r.chance(3)) nd.Annotations[GetUpgradeDriverWaitForSafeLoadAnnotationKey()] = "true";
    if (r.chance(5)) nd.Annotations[GetUpgradeInitialStateAnnotationKey()] = "true";
    if (r.chance(4)) nd.Conditions.push_back({"Ready", "False"});
    w.podObjs.emplace_back();
    Pod& p = w.podObjs.back();
    p.Name = "pod-" + std::to_string(i);
    p.ResourceVersion = "1";
    p.NodeName = nd.Name;
    p.Labels[PodControllerRevisionHashLabelKey] = r.chance(50) ? "test-hash-12345" : "test-hash-outdated";
    p.Phase = r.chance(90) ? "Running" : "Pending";
    p.ContainerStatuses = {{r.chance(85), (int)(r.next() % 14)}};
    p.DeletionTimestampSet = r.chance(3);
  }
}

// what a cluster does between two reconciles; identical on both worlds (same seed, same actuator records)
inline void evolve(World& w, Lcg r) {
  auto bump = [](Node* n) { VersionedProvider::bump(n); };
  for (Node* n : w.cordon.cordoned) { n->Unschedulable = true; bump(n); }
  for (Node* n : w.cordon.uncordoned) { n->Unschedulable = false; bump(n); }
  w.cordon.cordoned.clear(); w.cordon.uncordoned.clear();
  for (Pod* p : w.pods.restarted) {  // the DaemonSet controller recreates the pod at the current revision
    p->Labels[PodControllerRevisionHashLabelKey] = "test-hash-12345";
    p->Phase = "Running"; p->ContainerStatuses = {{true, 0}}; p->DeletionTimestampSet = false;
    p->ResourceVersion = std::to_string(std::stoll(p->ResourceVersion) + 1);
  }
  w.pods.restarted.clear();
  for (size_t i = 0; i < w.nodes.size(); i++) {
    Node& nd = w.nodes[i];
    if (nd.Name.empty()) continue;
    if (r.chance(2)) { nd.Labels[GetUpgradeSkipNodeLabelKey()] = r.chance(50) ? "true" : "false"; bump(&nd); }
    if (r.chance(1)) { nd.Annotations[GetUpgradeRequestedAnnotationKey()] = "true"; bump(&nd); }
    if (r.chance(1)) { w.podObjs[i].ContainerStatuses = {{r.chance(50), (int)(r.next() % 14)}}; w.podObjs[i].ResourceVersion = std::to_string(std::stoll(w.podObjs[i].ResourceVersion) + 1); }
    if (r.chance(1) && i % 5 == 4) nd.Name.clear();  // the node leaves the cluster
  }
}

inline std::string image(const World& w) {
  std::string s;
  for (const Node& n : w.nodes) {
    s += n.Name + "{" + getNodeUpgradeState(&n) + (n.Unschedulable ? ",U" : "");
    for (const auto& kv : n.Annotations) s += "," + kv.first + "=" + kv.second;
    s += "}";
  }
  return s;
}
inline std::string names(const std::vector<Node*>& v) { std::string s; for (auto* n : v) s += n->Name + ","; return s; }
inline std::string pnames(const std::vector<Pod*>& v) { std::string s; for (auto* p : v) s += p->Name + ","; return s; }

using WorldApplyFn = std::function<Error(World&, const DriverUpgradePolicySpec*)>;

inline void run_incremental(Runner& R, const MakeFn& makeFull, const WorldApplyFn& applyFull, const MakeFn& makeIncr,
                            const WorldApplyFn& applyIncr, int n_nodes) {
  SetDriverName("gpu");
  R.it("ApplyStateIncremental == ApplyState over a multi-reconcile rollout (resourceVersion-keyed encode cache)", [&] {
    World a, b;
    a.m = makeFull({}); a.wire();
    b.m = makeIncr({}); b.wire();
    populate(a, n_nodes, 42); populate(b, n_nodes, 42);
    DriverUpgradePolicySpec p;
    p.AutoUpgrade = true;
    p.MaxParallelUpgrades = 9;
    p.MaxUnavailable = IntOrString::FromString("30%");
    p.DrainSpec = upgrade::DrainSpec{};
    p.DrainSpec->Enable = true;
    for (int rec = 0; rec < 14; rec++) {
      if (rec == 6) p.MaxParallelUpgrades = 0;                           // the policy changes mid-rollout
      if (rec == 9) p.MaxUnavailable = IntOrString::FromString("50%");
      a.snapshot(); b.snapshot();
      const Error ea = applyFull(a, &p), eb = applyIncr(b, &p);
      EXPECT(R, ea.has_value() == eb.has_value());
      EXPECT(R, image(a) == image(b));
      EXPECT(R, names(a.cordon.cordoned) == names(b.cordon.cordoned));
      EXPECT(R, names(a.cordon.uncordoned) == names(b.cordon.uncordoned));
      EXPECT(R, pnames(a.pods.restarted) == pnames(b.pods.restarted));
      EXPECT(R, a.drain.calls == b.drain.calls && a.pods.evictionCalls == b.pods.evictionCalls && a.pods.waitCalls == b.pods.waitCalls);
      if (R.failed_here) { std::printf("    (reconcile %d)\n", rec); break; }
      evolve(a, Lcg{1000u + (uint64_t)rec}); evolve(b, Lcg{1000u + (uint64_t)rec});
    }
    const auto& st = b.m->Stats();
    std::printf("    incremental: %lld reconciles, %lld full uploads, %lld entries encoded, %lld reused, %lld outputs received\n",
                (long long)st.reconciles, (long long)st.full_uploads, (long long)st.encoded, (long long)st.reused, (long long)st.outputs_received);
    EXPECT(R, st.reused > st.encoded);      // most entries are served from the cache
    EXPECT(R, st.full_uploads <= 2);        // the first reconcile (and at most one re-ordering)
  });
}

}  // namespace spec
