"""The C++ mirror of the reference's manager interface (k8s-operator-libs_b200/host) against the reference's own
ApplyState / BuildState specs, restated in tests/host/upgrade_state_spec.hpp (one It() per Go It())."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
    return os.path.join(ROOT, "tests", "host", "_build")


def _run(exe):
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out
    ok = [line for line in out.splitlines() if line.startswith("ok ")]
    assert len(ok) >= 40, out
    assert "not ok" not in out, out
    return out


def test_host_halves_encode_replay_cpu():
    """Encode -> oracle (checker) -> Replay: the host logic on its own, no GPU."""
    _run(os.path.join(_build(), "host_logic_test"))


@pytest.mark.gpu
def test_reference_specs_on_gpu():
    """The same specs through ClusterUpgradeStateManagerImpl::ApplyState (C ABI, B200 kernel) + BuildState."""
    out = _run(os.path.join(_build(), "upgrade_state_test"))
    assert "BuildState should process running daemonset pods" in out
