// Host-side encoder throughput ("in production it dominates"): ClusterUpgradeState -> struct of arrays,
// on one host thread, for a cluster whose node objects look like real ones (two dozen labels, a few annotations).
// Needs no GPU:  g++ -O2 -std=c++17 -I. scripts/micro/encode_bench.cpp -Lk8s-operator-libs_b200 -lust_host -lust \
//                    -Wl,-rpath,$PWD/k8s-operator-libs_b200 -o /tmp/encode_bench && /tmp/encode_bench 1000000 [threads]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>

#include "tests/host/mocks.hpp"

using namespace upgrade;
using clk = std::chrono::steady_clock;

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  std::mt19937_64 rng(7);
  std::vector<std::unique_ptr<Node>> nodes;
  std::vector<std::unique_ptr<Pod>> pods;
  std::vector<DaemonSet> dss(4);
  for (int d = 0; d < 4; d++) { dss[d].Name = "driver-ds-" + std::to_string(d); dss[d].UID = "uid-" + std::to_string(d); }
  ClusterUpgradeState st = NewClusterUpgradeState();
  const char* states[] = {"", UpgradeStateUpgradeRequired, UpgradeStateCordonRequired, UpgradeStateWaitForJobsRequired,
                          UpgradeStatePodDeletionRequired, UpgradeStateDrainRequired, UpgradeStatePodRestartRequired,
                          UpgradeStateValidationRequired, UpgradeStateUncordonRequired, UpgradeStateDone, UpgradeStateFailed};
  const int weight[] = {5, 35, 5, 5, 5, 5, 10, 2, 5, 20, 3};
  for (long i = 0; i < n; i++) {
    auto node = std::make_unique<Node>();
    node->Name = "node-" + std::to_string(i);
    node->ResourceVersion = std::to_string(1000 + i);
    for (int l = 0; l < 24; l++) node->Labels["topology.example.com/label-" + std::to_string(l)] = "value-" + std::to_string((i + l) % 7);
    for (int a = 0; a < 4; a++) node->Annotations["node.example.com/annotation-" + std::to_string(a)] = "x";
    int r = (int)(rng() % 100), s = 0;
    while (r >= weight[s]) r -= weight[s++];
    if (*states[s]) node->Labels[GetUpgradeStateLabelKey()] = states[s];
    node->Unschedulable = rng() % 10 == 0;
    node->Conditions.push_back({"Ready", rng() % 50 == 0 ? "False" : "True"});
    if (rng() % 20 == 0) node->Annotations[GetUpgradeInitialStateAnnotationKey()] = "true";
    auto pod = std::make_unique<Pod>();
    pod->Name = "driver-" + std::to_string(i);
    pod->NodeName = node->Name;
    pod->ResourceVersion = std::to_string(5000 + i);
    pod->Phase = rng() % 20 ? "Running" : "Pending";
    pod->Labels[PodControllerRevisionHashLabelKey] = rng() % 2 ? "test-hash-12345" : "old-hash-6789";
    pod->ContainerStatuses.push_back({rng() % 10 != 0, 0});
    const int d = (int)(rng() % 4);
    pod->OwnerReferences.push_back({"DaemonSet", dss[d].Name, dss[d].UID});
    auto ns = std::make_unique<NodeUpgradeState>();
    ns->Node = node.get(); ns->DriverPod = pod.get(); ns->DriverDaemonSet = &dss[d]; ns->ListIndex = i;
    st.NodeStates[states[s]].push_back(ns.get());
    st.owned.push_back(std::move(ns));
    nodes.push_back(std::move(node)); pods.push_back(std::move(pod));
  }
  StateOptions so;
  so.EncodeThreads = argc > 2 ? atoi(argv[2]) : 1;
  auto m = ClusterUpgradeStateManagerImpl::NewDetached(so);
  mocks::NodeUpgradeStateProviderMock provider; mocks::CordonManagerMock cordon; mocks::DrainManagerMock drain;
  mocks::PodManagerMock podm; mocks::ValidationManagerMock valid; mocks::SafeDriverLoadManagerImpl safe(&provider);
  m->NodeUpgradeStateProvider = &provider; m->CordonManager = &cordon; m->DrainManager = &drain; m->PodManager = &podm;
  m->ValidationManager = &valid; m->SafeDriverLoadManager = &safe;
  DriverUpgradePolicySpec pol; pol.AutoUpgrade = true; pol.MaxParallelUpgrades = 100; pol.MaxUnavailable = IntOrString::FromString("25%");
  EncodedSnapshot enc;
  double best = 1e30;
  for (int rep = 0; rep < 3; rep++) {
    auto t0 = clk::now();
    if (auto e = m->Encode(st, pol, &enc)) { std::printf("encode error: %s\n", e->c_str()); return 1; }
    const double s = std::chrono::duration<double>(clk::now() - t0).count();
    if (s < best) best = s;
  }
  std::printf("Encode (%d thread(s)): %ld nodes in %.3f s = %.2f M nodes/s (%.0f ns per node), %zu entries\n", so.EncodeThreads, n, best, n / best / 1e6, best / n * 1e9, enc.entries.size());
  return 0;
}
