#!/usr/bin/env python
"""bench.py — node state-transitions/s of ApplyState on B200 (BASELINE.json metric).

A "step" is one ApplyState pass over one synthetic cluster snapshot. Workload at N=1: BASELINE config C3
(10 M nodes, MaxParallelUpgrades=100, MaxUnavailable=25%). At N>1 every rank holds one 10 M-node contiguous
shard of an N x 10 M-node cluster (config C5 at N=8): weak scaling, one exchange of the constraint counters
per step.

  value      device-resident inputs, CUDA-event time of K steps, max over ranks, whole-job nodes/s
  e2e        the same step through the host-pointer C ABI, pinned host arrays in, H2D + kernel + D2H inside the
             timed region. N=1: ust_apply_state_packed (uint16 revisions / int8 DaemonSet indices on the host side,
             8 B/node up; --e2e-format wide = ust_apply_state, 13 B/node up); N>1: ust_apply_state
  e2e_delta  (N=1, informative) ust_apply_state_delta: the snapshot stays resident, 1 % of the nodes are
             re-encoded and uploaded per step, everything is evaluated, all outputs come back
  roofline   dominant kernel (ust_fused_kernel): 16 algorithmic bytes per node / its CUDA-event duration,
             against the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline / --impl reference   the oracle's reference-shaped restatement of the Go loop (1 thread —
             the reference's ApplyState is strictly sequential), bounded sample of the same workload
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "k8s-operator-libs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

BYTES_PER_NODE = 16  # state 1 + flags 4 + pod_rev 4 + ds_idx 4 read, next_state 1 + actions 2 written
SHARD_NODES = 10_000_000
CPU_SAMPLE_NODES = 1_000_000
L2_BYTES = 126 * 1024 * 1024


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    ncu capture (profiles/rNN_ncu_traffic.json); None when there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(policy, soa, reps):
    """Reference-shaped oracle (ApplyState only, objects built untimed) on one host thread. nodes/s."""
    import helpers
    n = int(soa["state"].shape[0])
    sec = helpers.oracle().ust_oracle_time_apply_state(
        C.c_int(0), C.byref(policy), C.c_int64(n), soa["state"].ctypes.data_as(C.c_void_p),
        soa["flags"].ctypes.data_as(C.c_void_p), soa["pod_rev"].ctypes.data_as(C.c_void_p),
        soa["ds_idx"].ctypes.data_as(C.c_void_p), C.c_int32(int(soa["ds_rev"].shape[0])),
        soa["ds_rev"].ctypes.data_as(C.c_void_p), None, C.c_int(reps))
    return n / sec, sec


def run_reference(args, rank, world):
    """--impl reference: the reference's own (sequential, CPU) ApplyState, represented by the oracle's
    reference-shaped restatement — Go is not installed here or on the GPU box."""
    if rank != 0:
        return
    from ust import synth
    cfg = synth.CONFIGS["C3" if world == 1 else "C5"]
    pol = synth.config_policy("C3")
    soa = synth.make_nodes(CPU_SAMPLE_NODES, cfg["seed"])
    cpu_reference_rate(pol, soa, max(1, min(args.warmup, 1)))
    t0 = time.time()
    rate, sec = cpu_reference_rate(pol, soa, max(1, args.steps))
    sample = f"first {CPU_SAMPLE_NODES} nodes of the workload, {max(1, args.steps)} ApplyState passes, median"
    line = {
        "impl": "reference", "metric": "node state-transitions/sec", "value": rate, "unit": "nodes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32", "data": "synthetic",
        "config": {"workload": workload_name(world), "sample_nodes": CPU_SAMPLE_NODES},
        "cpu_baseline": {"value": rate, "unit": "nodes/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def workload_name(world):
    if world == 1:
        return "C3: 10M-node synthetic cluster, MaxParallelUpgrades=100, MaxUnavailable=25%, 1xB200"
    return (f"C5-style: {world}x10M-node contiguous shards ({world * 10}M nodes), MaxParallelUpgrades=100, "
            f"MaxUnavailable=25%, one exchange of the constraint counters per ApplyState")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ust", choices=["ust", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--nodes", type=int, default=SHARD_NODES, help="nodes per GPU (default: the BASELINE workload)")
    ap.add_argument("--sets", type=int, default=16, help="rotating input/output buffer sets (L2 defeat)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"], help="N>1: counter exchange mechanism")
    ap.add_argument("--maxpar", type=int, default=None, help="tuning: override MaxParallelUpgrades")
    ap.add_argument("--maxunav", default=None, help="tuning: override MaxUnavailable ('nil', int or 'NN%%')")
    ap.add_argument("--quick", action="store_true", help="tuning: device-resident timing only (no e2e / cpu baseline)")
    ap.add_argument("--e2e-format", default="packed", choices=["wide", "packed"],
                    help="host format of the e2e leg: wide = ust_apply_state (int32 pod_rev / ds_idx), packed = "
                         "ust_apply_state_packed (uint16 / int8: 8 instead of 13 bytes per node over PCIe)")
    ap.add_argument("--pods", action="store_true", help="tuning (with --quick): the C4 workload - CSR pod lists, pod deletion and drain enabled")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from ust import abi, lib as ustlib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libust.so has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    warmup = max(args.warmup, 3)
    n = args.nodes
    cfg = synth.CONFIGS["C3" if world == 1 else "C5"]
    pol = synth.config_policy("C3")
    if args.maxpar is not None or args.maxunav is not None:
        mu = {"nil": None}.get(args.maxunav, args.maxunav) if args.maxunav is not None else "25%"
        if isinstance(mu, str) and mu.isdigit():
            mu = int(mu)
        pol = abi.make_policy(max_parallel_upgrades=args.maxpar if args.maxpar is not None else 100, max_unavailable=mu)
    if args.pods:
        if not args.quick:
            raise SystemExit("--pods is a tuning option: use it with --quick")
        cfg = synth.CONFIGS["C4"]
        pol = synth.config_policy("C4")
    soa = synth.make_nodes(n, cfg["seed"], start=rank * n)
    n_ds = int(soa["ds_rev"].shape[0])
    pods_dev = pods_struct = None
    if args.pods:
        pods = synth.make_pods_blocked(n, cfg["seed"], start=rank * n)
        pods_dev = {k: torch.from_numpy(v).to(dev) for k, v in pods.items()}
        pods_struct = abi.Pods(pods_dev["pod_off"].data_ptr(), pods_dev["pod_flags"].data_ptr(), int(pods["pod_flags"].shape[0]))

    h = ustlib.Handle(local_rank)
    if world > 1:
        uid = [ustlib.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(rank, world, uid[0])
        exchange = "ncclAllReduce of 42 int64 lanes between two kernels"
        if args.exchange == "fused":
            # every rank must end up in the same mode: agree on whether all of them mapped their peers
            ok = torch.ones(1, device=dev)
            try:
                h.comm_set_mode(1)
            except ustlib.UstError:
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() > 0:
                exchange = "fused in-kernel NVLink mailbox exchange (CUDA IPC peer memory), one kernel per rank"
            else:
                h.comm_set_mode(0)

    # Rotating buffer sets: a step never finds more than 126 MB / (SETS x 160 MB) of its inputs in L2.
    # (Two sets are NOT enough: the streaming loads are evict-first, so L2 keeps a fixed ~126 MB subset of
    # the 320 MB alive and half of every step would be L2 hits — measured in round 1, profiles/README.md.)
    SETS = max(2, args.sets)
    bufs = []
    for _ in range(SETS):
        d = {k: torch.from_numpy(v).to(dev) for k, v in soa.items()}
        d["next"] = torch.empty(n, dtype=torch.uint8, device=dev)
        d["actions"] = torch.empty(n, dtype=torch.int16, device=dev)
        d["outcome"] = torch.empty(n, dtype=torch.uint8, device=dev) if args.pods else None
        bufs.append(d)
    counters = torch.zeros(C.sizeof(abi.Counters) // 8, dtype=torch.int64, device=dev)
    # a dedicated stream: torch's default stream has handle 0, which the C ABI reads as "use the handle's own
    # stream" — events recorded on torch's stream would then not bracket the kernels (round-1 lesson)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # pre-bound ctypes arguments per buffer set: the launch loop must not be the bottleneck (a step is ~40 us)
    fn = ustlib.load().ust_apply_state_device
    pol_p = C.c_void_p(C.addressof(pol))
    cnt_p = C.c_void_p(counters.data_ptr())
    st_p = C.c_void_p(stream)
    bound = []
    for b in bufs:
        bound.append((h._h, pol_p, C.c_int64(n), C.c_void_p(b["state"].data_ptr()), C.c_void_p(b["flags"].data_ptr()),
                      C.c_void_p(b["pod_rev"].data_ptr()), C.c_void_p(b["ds_idx"].data_ptr()), C.c_int32(n_ds),
                      C.c_void_p(b["ds_rev"].data_ptr()), C.byref(pods_struct) if pods_struct is not None else None,
                      C.c_void_p(b["next"].data_ptr()), C.c_void_p(b["actions"].data_ptr()),
                      C.c_void_p(b["outcome"].data_ptr()) if b["outcome"] is not None else None, cnt_p, st_p))

    def step(i):
        rc = fn(*bound[i % SETS])
        if rc:
            raise ustlib.UstError(rc, h.last_error())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def blocker():
        # ~1.5 ms of GPU spinning: the host queues the timed launches behind it, so the CUDA events bracket
        # back-to-back GPU execution rather than the Python launch rate
        torch.cuda._sleep(3_000_000)

    for i in range(warmup):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = h.launch_count()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    blocker()
    e_start.record()
    for i in range(args.steps):
        step(i)
    e_end.record()
    barrier()
    t_wall = time.time() - t_wall0
    launches = h.launch_count() - launches0
    total_ms = e_start.elapsed_time(e_end)
    # dominant-kernel duration: per-step event pairs over a second, shorter run of the same steps
    ksteps = min(args.steps, 20)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ksteps)]
    blocker()
    for i in range(ksteps):
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    barrier()
    per_step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    clocks = sampler.stop()

    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # verify the outputs of the timed buffers (rank-local properties; full parity lives in tests/)
    cnt = abi.Counters.from_buffer_copy(counters.cpu().numpy().tobytes()).as_dict()
    nxt = bufs[0]["next"].cpu().numpy()
    code = soa["state"] & 15
    assert cnt["error_code"] == 0
    assert cnt["total_managed"] == world * n or world > 1
    assert np.array_equal(nxt[code == 2], np.full(int((code == 2).sum()), 3, np.uint8)), "cordon-required -> wait-for-jobs"
    if args.quick and rank == 0:
        print("counters:", {k: cnt[k] for k in ("candidates", "upgrades_available", "max_unavailable")},
              "redone chunks:", abi.Counters.from_buffer_copy(counters.cpu().numpy().tobytes()).reserved[0], flush=True)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        # one step == one launch of the dominant kernel: its average duration over the timed region; the
        # per-step event pairs (each pair adds ~2 us of event latency) are reported alongside
        kern_ms = total_ms / args.steps
        achieved = BYTES_PER_NODE * n / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "node state-transitions/sec", "value": value, "unit": "nodes/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32", "data": "synthetic",
            "config": {"workload": workload_name(world), "nodes_per_gpu": n, "bytes_per_node": BYTES_PER_NODE,
                       "l2": f"inputs larger than L2: {SETS} rotating buffer sets of {BYTES_PER_NODE * n / 1e6:.0f} MB each "
                             f"({SETS * BYTES_PER_NODE * n / 1e9:.2f} GB vs 126 MB L2, at most {100 * 126e6 / (SETS * BYTES_PER_NODE * n):.0f}% of a step can hit)",
                       "exchange": "none" if world == 1 else exchange},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic() if (world == 1 and n == SHARD_NODES) else None, "peak_source": peak_src,
                         "kernel": "ust_stream_kernel",
                         "kernel_ms": kern_ms, "kernel_ms_event_pairs_median": float(np.median(per_step_ms)),
                         "frac_of_8TBs": achieved / 8000.0},
            "clocks": clocks, "gpu_launches": int(launches), "wall_s_timed_region": t_wall,
            "counters": {k: cnt[k] for k in ("total_managed", "in_progress", "unavailable", "max_unavailable", "upgrades_available")},
        }

    if args.quick:
        if rank == 0 and os.environ.get("UST_STAMPS"):
            g = min(int(os.environ["UST_STAMPS"]), 148)
            st = (C.c_uint64 * (4 * g + 4))()
            ustlib.load().ust_debug_stamps(h._h, st, g)
            a = np.array(st, dtype=np.int64)
            v = a[4 * g:]
            a = a[:4 * g].reshape(g, 4)
            a = a[a[:, 0] > a[:, 0].max() - 1_000_000]   # CTAs of the last launch only (a small snapshot uses fewer)
            t0 = a[:, 0].min()
            rel = (a - t0) / 1e3
            v = (v - t0) / 1e3
            print("stamps us: entry[min,max]=%.1f,%.1f first_tile[min,med,max]=%.1f,%.1f,%.1f stream_end[min,med,max]=%.1f,%.1f,%.1f "
                  "exit[max]=%.1f | verify kernel CTA 0: woken %.1f vector %.1f decided %.1f redo done %.1f" % (
                      rel[:, 0].min(), rel[:, 0].max(), rel[:, 1].min(), np.median(rel[:, 1]), rel[:, 1].max(),
                      rel[:, 2].min(), np.median(rel[:, 2]), rel[:, 2].max(), rel[:, 3].max(), v[0], v[1], v[2], v[3]), flush=True)
        if rank == 0:
            print(json.dumps({k: line[k] for k in ("value", "ms_per_step", "roofline", "clocks")}), flush=True)
        h.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: host-pointer C ABI with pinned host buffers, H2D + kernel + D2H inside the timed region ----
    host = {k: ustlib.pinned_array(v.shape, v.dtype) for k, v in soa.items()}
    for k in soa:
        host[k][...] = soa[k]
    out = (ustlib.pinned_array(n, np.uint8), ustlib.pinned_array(n, np.uint16), None)
    packed = args.e2e_format == "packed" and world == 1   # N > 1 keeps the int32 host format
    if packed:
        assert soa["pod_rev"].min() >= 0 and soa["pod_rev"].max() < 65536 and soa["ds_idx"].min() >= -128 and n_ds <= 127
        pk = (ustlib.pinned_array(n, np.uint16), ustlib.pinned_array(n, np.int8))
        pk[0][...] = soa["pod_rev"]
        pk[1][...] = soa["ds_idx"]

    def e2e_step():
        if packed:
            h.apply_state_packed(pol, host, want_outcome=False, out=out, check=True, packed=pk)
        else:
            h.apply_state(pol, host, want_outcome=False, out=out, check=True)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.time()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.time() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.e2e_steps / float(te.item())
    assert np.array_equal(out[0], nxt), "e2e result differs from the device-resident result"

    # ---- the same through the delta entry point: the snapshot stays resident, 1 % of the nodes are re-encoded and
    # re-uploaded per step (a reconcile that watches resourceVersions), the whole snapshot is evaluated, all outputs
    # come back. Informative only - `e2e` above (full upload every step) is the headline.
    delta = None
    if world == 1:
        rng = np.random.default_rng(7)
        m = max(1, n // 100)
        steps_d = max(3, args.e2e_steps)
        deltas = []
        for _ in range(steps_d + 1):
            idx = rng.choice(n, size=m, replace=False).astype(np.int64)
            src = rng.integers(0, n, size=m)
            deltas.append((idx, {k: np.ascontiguousarray(soa[k][src]) for k in ("state", "flags", "pod_rev", "ds_idx")}))
        h.apply_state_delta(pol, n, deltas[0][0], deltas[0][1], soa["ds_rev"], want_outcome=False, out=out)
        torch.cuda.synchronize()
        t0 = time.time()
        for idx, ch in deltas[1:]:
            rc = h.apply_state_delta(pol, n, idx, ch, soa["ds_rev"], want_outcome=False, out=out)[0]
            assert rc == 0, h.last_error()
        torch.cuda.synchronize()
        d_s = time.time() - t0
        delta = {"value": n * steps_d / d_s, "unit": "nodes/s", "ms_per_step": d_s / steps_d * 1e3, "changed_nodes_per_step": m,
                 "h2d_bytes_per_step": 21 * m + 4 * n_ds, "d2h_bytes_per_step": 3 * n + C.sizeof(abi.Counters), "steps": steps_d}

    if rank == 0:
        if delta is not None:
            line["e2e_delta"] = delta
        line["e2e"] = {"value": e2e_value, "unit": "nodes/s", "h2d_bytes_per_step": (8 if packed else 13) * n + 4 * n_ds,
                       "entry_point": "ust_apply_state_packed" if packed else "ust_apply_state",
                       "d2h_bytes_per_step": 3 * n + C.sizeof(abi.Counters), "steps": args.e2e_steps,
                       "ms_per_step": float(te.item()) / args.e2e_steps * 1e3}
        # ---- CPU baseline beside it: bounded sample, 1 thread (the reference loop is sequential) ----
        m = min(CPU_SAMPLE_NODES, n)
        sample = {k: (v[:m].copy() if k != "ds_rev" else v) for k, v in soa.items()}
        rate, sec = cpu_reference_rate(pol, sample, 3)
        line["cpu_baseline"] = {"value": rate, "unit": "nodes/s", "cores": 1, "kind": "port",
                                "sample": f"first {m} nodes of the workload, ApplyState only, median of 3 passes "
                                          f"({sec:.2f} s each), reference-shaped oracle (Go unavailable)"}
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
