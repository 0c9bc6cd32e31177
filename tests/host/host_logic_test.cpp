// Host halves only (no GPU): Encode -> ORACLE -> Replay. The oracle stands in for the kernel here purely as a
// checker of the encoder / replayer logic; the product's ApplyState never does this (it fails without a device).
#include <cstring>

#include "incremental_spec.hpp"

extern "C" int ust_oracle_apply_state(int variant, const ust_policy* policy, int64_t n, const uint8_t* state,
                                      const uint32_t* flags, const int32_t* pod_rev, const int32_t* ds_idx, int32_t n_ds,
                                      const int32_t* ds_rev, const ust_pods* pods, uint8_t* next_state, uint16_t* actions,
                                      uint8_t* actuator_outcome, ust_counters* out);

int main() {
  mocks::Runner R;
  spec::MakeFn make = [](upgrade::StateOptions o) { return upgrade::ClusterUpgradeStateManagerImpl::NewDetached(o); };
  spec::ApplyFn apply = [](spec::Env& e, upgrade::ClusterUpgradeState* s, const upgrade::DriverUpgradePolicySpec* p) -> upgrade::Error {
    if (s == nullptr) return upgrade::Errorf("currentState should not be empty");
    if (p == nullptr || !p->AutoUpgrade) return std::nullopt;
    upgrade::EncodedSnapshot enc;
    if (auto err = e.m->Encode(*s, *p, &enc)) return err;
    const size_t n = enc.entries.size();
    std::vector<uint8_t> next(n + 1);
    std::vector<uint16_t> actions(n + 1);
    enc.state.push_back(0); enc.flags.push_back(0); enc.pod_rev.push_back(0); enc.ds_idx.push_back(0); enc.ds_rev.push_back(0);
    ust_counters c;
    const int rc = ust_oracle_apply_state(0, &enc.policy, (int64_t)n, enc.state.data(), enc.flags.data(), enc.pod_rev.data(),
                                          enc.ds_idx.data(), (int32_t)enc.ds_rev.size() - 1, enc.ds_rev.data(), nullptr,
                                          next.data(), actions.data(), nullptr, &c);
    return e.m->Replay(enc, *p, next.data(), actions.data(), rc, c);
  };
  spec::run(R, make, apply);
  // the detached manager must refuse to decide anything by itself
  R.it("a manager without a device refuses ApplyState (no CPU path)", [&] {
    auto m = upgrade::ClusterUpgradeStateManagerImpl::NewDetached({});
    upgrade::ClusterUpgradeState s;
    upgrade::DriverUpgradePolicySpec p;
    p.AutoUpgrade = true;
    auto err = m->ApplyState(&s, &p);
    EXPECT(R, err.has_value());
  });
  // the incremental path with the oracle behind the cache (the product evaluates the cache on the device)
  struct OracleBacked : upgrade::ClusterUpgradeStateManagerImpl {
    int EvaluateCached(const ust_policy& policy, bool, const std::vector<int64_t>&, Cache* cache, ust_counters* c) override {
      Cache& k = *cache;
      const size_t n = k.slots.size();
      k.next.assign(n + 1, 0);
      k.actions.assign(n + 1, 0);
      std::vector<uint8_t> st = k.state; st.push_back(0);
      std::vector<uint32_t> fl = k.flags; fl.push_back(0);
      std::vector<int32_t> rv = k.pod_rev, di = k.ds_idx, dr = k.ds_rev;
      rv.push_back(0); di.push_back(0); dr.push_back(0);
      return ust_oracle_apply_state(0, &policy, (int64_t)n, st.data(), fl.data(), rv.data(), di.data(), (int32_t)k.ds_rev.size(),
                                    dr.data(), nullptr, k.next.data(), k.actions.data(), nullptr, c);
    }
  };
  spec::MakeFn makeIncr = [](upgrade::StateOptions) { return std::unique_ptr<upgrade::ClusterUpgradeStateManagerImpl>(new OracleBacked()); };
  spec::WorldApplyFn wfull = [&](spec::World& w, const upgrade::DriverUpgradePolicySpec* p) -> upgrade::Error {
    upgrade::EncodedSnapshot enc;
    if (auto err = w.m->Encode(w.state, *p, &enc)) return err;
    const size_t n = enc.entries.size();
    std::vector<uint8_t> next(n + 1);
    std::vector<uint16_t> actions(n + 1);
    enc.state.push_back(0); enc.flags.push_back(0); enc.pod_rev.push_back(0); enc.ds_idx.push_back(0); enc.ds_rev.push_back(0);
    ust_counters c;
    const int rc = ust_oracle_apply_state(0, &enc.policy, (int64_t)n, enc.state.data(), enc.flags.data(), enc.pod_rev.data(),
                                          enc.ds_idx.data(), (int32_t)enc.ds_rev.size() - 1, enc.ds_rev.data(), nullptr,
                                          next.data(), actions.data(), nullptr, &c);
    return w.m->Replay(enc, *p, next.data(), actions.data(), rc, c);
  };
  spec::WorldApplyFn wincr = [](spec::World& w, const upgrade::DriverUpgradePolicySpec* p) { return w.m->ApplyStateIncremental(&w.state, p); };
  spec::run_incremental(R, make, wfull, makeIncr, wincr, 400);
  // Encode with several host threads (StateOptions::EncodeThreads) against the single-threaded walk
  R.it("Encode on several host threads == Encode on one (revision ids up to renaming)", [&] {
    using namespace upgrade;
    SetDriverName("gpu");
    spec::World a, b;
    StateOptions one, many;
    many.EncodeThreads = 5;
    a.m = ClusterUpgradeStateManagerImpl::NewDetached(one); a.wire();
    b.m = ClusterUpgradeStateManagerImpl::NewDetached(many); b.wire();
    spec::populate(a, 30000, 77); spec::populate(b, 30000, 77);
    // revision hashes no DaemonSet has, different ones in different parts of the node list
    for (size_t i = 0; i < a.podObjs.size(); i += 7) {
      const std::string h = "stale-" + std::to_string(i / 9000) + "-" + std::to_string(i % 3);
      a.podObjs[i].Labels[PodControllerRevisionHashLabelKey] = h;
      b.podObjs[i].Labels[PodControllerRevisionHashLabelKey] = h;
    }
    a.snapshot(); b.snapshot();
    DriverUpgradePolicySpec p;
    p.AutoUpgrade = true;
    EncodedSnapshot ea, eb;
    EXPECT(R, !a.m->Encode(a.state, p, &ea).has_value());
    EXPECT(R, !b.m->Encode(b.state, p, &eb).has_value());
    EXPECT(R, ea.entries.size() == eb.entries.size() && ea.entries.size() > 25000);
    EXPECT(R, ea.state == eb.state && ea.flags == eb.flags && ea.ds_idx == eb.ds_idx && ea.ds_rev == eb.ds_rev);
    EXPECT(R, ea.deferred == eb.deferred);
    bool same_nodes = true, renaming = true;
    std::map<int32_t, int32_t> fwd, back;
    for (size_t i = 0; i < ea.entries.size() && i < eb.entries.size(); i++) {
      same_nodes = same_nodes && ea.entries[i]->Node->Name == eb.entries[i]->Node->Name;
      const int32_t x = ea.pod_rev[i], y = eb.pod_rev[i];
      auto f = fwd.emplace(x, y).first;
      auto g = back.emplace(y, x).first;
      renaming = renaming && f->second == y && g->second == x && (x == 0) == (y == 0);
      if (ea.ds_idx[i] >= 0) renaming = renaming && (x == ea.ds_rev[(size_t)ea.ds_idx[i]]) == (y == eb.ds_rev[(size_t)eb.ds_idx[i]]);
    }
    EXPECT(R, same_nodes);
    EXPECT(R, renaming);
    EXPECT(R, fwd.size() >= 10);  // the stale hashes made it into the table
  });
  std::printf("# %d passed, %d failed\n", R.passed, R.failed);
  return R.failed == 0 ? 0 : 1;
}
