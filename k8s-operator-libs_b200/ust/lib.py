"""ctypes binding of libust.so (include/ust.h). Plumbing for tests and bench.py — the product is the .so."""
import ctypes as C
import os

import numpy as np

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("UST_LIB") or os.path.join(os.path.dirname(HERE), "libust.so")  # UST_LIB: tuning experiments only

_lib = None


class UstError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{abi.ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


def load():
    """Load libust.so. There is no fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise UstError(abi.UST_ERR_CUDA, f"{SO_PATH} not built (python __graft_entry__.py build)")
        lib = C.CDLL(SO_PATH)
        lib.ust_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        lib.ust_destroy.argtypes = [C.c_void_p]
        lib.ust_destroy.restype = None
        lib.ust_last_error.argtypes = [C.c_void_p]
        lib.ust_last_error.restype = C.c_char_p
        lib.ust_create_error.restype = C.c_char_p
        lib.ust_launch_count.argtypes = [C.c_void_p]
        lib.ust_launch_count.restype = C.c_int64
        lib.ust_host_alloc.argtypes = [C.c_size_t]
        lib.ust_host_alloc.restype = C.c_void_p
        lib.ust_host_free.argtypes = [C.c_void_p]
        lib.ust_host_free.restype = None
        lib.ust_sync.argtypes = [C.c_void_p]
        lib.ust_stream.argtypes = [C.c_void_p]
        lib.ust_stream.restype = C.c_void_p
        apply_args = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ust_apply_state.argtypes = apply_args
        lib.ust_apply_state_device.argtypes = apply_args + [C.c_void_p]
        lib.ust_apply_state_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ust_apply_state_delta.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ust_apply_state_delta_sparse.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_void_p]
        lib.ust_fetch_outputs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ust_simulate_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ust_simulate_rollout_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p]
        lib.ust_build_state.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.ust_build_state_uids.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        lib.ust_get_unique_id.argtypes = [C.c_void_p]
        lib.ust_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.ust_comm_set_mode.argtypes = [C.c_void_p, C.c_int]
        lib.ust_table_entry.argtypes = [C.c_void_p, C.c_uint, C.c_uint32]
        lib.ust_table_entry.restype = C.c_uint32
        lib.ust_table_window_shift.argtypes = [C.c_uint]
        _lib = lib
    return _lib


EXPORTS = ["ust_abi_version", "ust_create", "ust_destroy", "ust_last_error", "ust_create_error", "ust_launch_count",
           "ust_host_alloc", "ust_host_free", "ust_apply_state", "ust_apply_state_device", "ust_stream", "ust_apply_state_packed", "ust_apply_state_delta", "ust_apply_state_delta_sparse", "ust_fetch_outputs", "ust_simulate_rollout", "ust_simulate_rollout_timed", "ust_sync",
           "ust_build_state", "ust_build_state_uids", "ust_get_unique_id", "ust_comm_init", "ust_comm_set_mode", "ust_table_entry",
           "ust_table_window_shift"]


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)  # raw address (e.g. torch tensor .data_ptr())


def pinned_array(shape, dtype):
    """numpy array backed by ust_host_alloc (page-locked) memory."""
    lib = load()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if not np.isscalar(shape) else int(shape)
    nbytes = max(n * dt.itemsize, 1)
    ptr = lib.ust_host_alloc(nbytes)
    if not ptr:
        raise UstError(abi.UST_ERR_CUDA, "ust_host_alloc failed")
    buf = (C.c_char * nbytes).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    _PINNED[arr.ctypes.data] = ptr
    return arr


_PINNED = {}


def free_pinned(arr):
    ptr = _PINNED.pop(arr.ctypes.data, None)
    if ptr:
        load().ust_host_free(ptr)


class Handle:
    """ust_handle wrapper. One per process per GPU."""

    def __init__(self, device=0):
        lib = load()
        h = C.c_void_p()
        rc = lib.ust_create(C.byref(h), device)
        if rc != 0:
            raise UstError(rc, lib.ust_create_error().decode())
        self._h = h
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ust_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self):
        return self._lib.ust_last_error(self._h).decode()

    def launch_count(self):
        return int(self._lib.ust_launch_count(self._h))

    def overlapped_calls(self):
        """diagnostics: how many ust_apply_state_device calls started without waiting for the previous call's tail"""
        fn = self._lib.ust_debug_relaxed_calls
        fn.restype = C.c_longlong
        fn.argtypes = [C.c_void_p]
        return int(fn(self._h))

    def stream(self):
        """cudaStream_t of the handle's own stream (as an int)."""
        return int(self._lib.ust_stream(self._h))

    def sync(self):
        rc = self._lib.ust_sync(self._h)
        if rc:
            raise UstError(rc, self.last_error())

    def apply_state(self, policy, soa, pods=None, want_outcome=True, out=None, check=False):
        """Host-array entry point (ust_apply_state). Returns (rc, next_state, actions, outcome, counters-dict)."""
        n = int(soa["state"].shape[0])
        if out is None:
            nxt = np.zeros(n, np.uint8)
            act = np.zeros(n, np.uint16)
            oc = np.full(n, 0xFF, np.uint8) if want_outcome else None
        else:
            nxt, act, oc = out
        cnt = abi.Counters()
        ps = None
        if pods is not None:
            off = np.ascontiguousarray(pods["pod_off"], dtype=np.int32)
            pf = np.ascontiguousarray(pods["pod_flags"], dtype=np.uint16)
            ps = abi.Pods(off.ctypes.data, pf.ctypes.data, int(pf.shape[0]))
        rc = self._lib.ust_apply_state(
            self._h, C.addressof(policy) if policy is not None else None, n, _p(soa["state"]), _p(soa["flags"]),
            _p(soa["pod_rev"]), _p(soa["ds_idx"]), int(soa["ds_rev"].shape[0]), _p(soa["ds_rev"]),
            C.addressof(ps) if ps is not None else None, _p(nxt), _p(act), _p(oc), C.addressof(cnt))
        if check and rc:
            raise UstError(rc, self.last_error())
        return rc, nxt, act, oc, cnt.as_dict()

    def apply_state_packed(self, policy, soa, want_outcome=True, out=None, check=False, packed=None):
        """ust_apply_state_packed: the same snapshot with pod_rev as uint16 and ds_idx as int8 on the host side.
        `packed` = (pod_rev16, ds_idx8) arrays to reuse (e.g. pinned); by default they are made from soa."""
        n = int(soa["state"].shape[0])
        if packed is None:
            assert n == 0 or (soa["pod_rev"].min() >= 0 and soa["pod_rev"].max() < 65536 and soa["ds_idx"].min() >= -128 and soa["ds_idx"].max() < 128)
            packed = (np.ascontiguousarray(soa["pod_rev"], dtype=np.uint16), np.ascontiguousarray(soa["ds_idx"], dtype=np.int8))
        if out is None:
            nxt = np.zeros(n, np.uint8)
            act = np.zeros(n, np.uint16)
            oc = np.full(n, 0xFF, np.uint8) if want_outcome else None
        else:
            nxt, act, oc = out
        cnt = abi.Counters()
        rc = self._lib.ust_apply_state_packed(
            self._h, C.addressof(policy) if policy is not None else None, n, _p(soa["state"]), _p(soa["flags"]), _p(packed[0]),
            _p(packed[1]), int(soa["ds_rev"].shape[0]), _p(soa["ds_rev"]), _p(nxt), _p(act), _p(oc), C.addressof(cnt))
        if check and rc:
            raise UstError(rc, self.last_error())
        return rc, nxt, act, oc, cnt.as_dict()

    def apply_state_delta(self, policy, n, idx, changed, ds_rev, want_outcome=True, out=None):
        """ust_apply_state_delta: overwrite nodes `idx` of the resident snapshot (n nodes) with `changed`
        (dict of state / flags / pod_rev / ds_idx arrays of len(idx)) and evaluate it again."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        ch = {"state": np.ascontiguousarray(changed["state"], dtype=np.uint8),
              "flags": np.ascontiguousarray(changed["flags"], dtype=np.uint32),
              "pod_rev": np.ascontiguousarray(changed["pod_rev"], dtype=np.int32),
              "ds_idx": np.ascontiguousarray(changed["ds_idx"], dtype=np.int32)}
        ds_rev = np.ascontiguousarray(ds_rev, dtype=np.int32)
        if out is None:
            nxt = np.zeros(n, np.uint8)
            act = np.zeros(n, np.uint16)
            oc = np.full(n, 0xFF, np.uint8) if want_outcome else None
        else:
            nxt, act, oc = out
        cnt = abi.Counters()
        rc = self._lib.ust_apply_state_delta(
            self._h, C.addressof(policy) if policy is not None else None, int(idx.shape[0]), _p(idx), _p(ch["state"]),
            _p(ch["flags"]), _p(ch["pod_rev"]), _p(ch["ds_idx"]), int(ds_rev.shape[0]), _p(ds_rev), _p(nxt), _p(act), _p(oc),
            C.addressof(cnt))
        return rc, nxt, act, oc, cnt.as_dict()

    def apply_state_delta_sparse(self, policy, idx, changed, ds_rev, max_out, out=None):
        """ust_apply_state_delta_sparse: like apply_state_delta, but only the outputs that differ from the previous call's
        come back. Returns (rc, n_out, out_idx, out_next, out_actions, counters-dict); the arrays hold n_out entries
        when n_out <= max_out."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        ch = {"state": np.ascontiguousarray(changed["state"], dtype=np.uint8),
              "flags": np.ascontiguousarray(changed["flags"], dtype=np.uint32),
              "pod_rev": np.ascontiguousarray(changed["pod_rev"], dtype=np.int32),
              "ds_idx": np.ascontiguousarray(changed["ds_idx"], dtype=np.int32)}
        ds_rev = np.ascontiguousarray(ds_rev, dtype=np.int32)
        if out is None:
            out = (np.zeros(max_out + 1, np.int64), np.zeros(max_out + 1, np.uint8), np.zeros(max_out + 1, np.uint16))
        n_out = C.c_int64(0)
        cnt = abi.Counters()
        rc = self._lib.ust_apply_state_delta_sparse(
            self._h, C.addressof(policy) if policy is not None else None, int(idx.shape[0]), _p(idx), _p(ch["state"]),
            _p(ch["flags"]), _p(ch["pod_rev"]), _p(ch["ds_idx"]), int(ds_rev.shape[0]), _p(ds_rev), C.c_int64(int(max_out)),
            _p(out[0]), _p(out[1]), _p(out[2]), C.addressof(n_out), C.addressof(cnt))
        return rc, int(n_out.value), out[0], out[1], out[2], cnt.as_dict()

    def fetch_outputs(self, n):
        nxt = np.zeros(n, np.uint8)
        act = np.zeros(n, np.uint16)
        rc = self._lib.ust_fetch_outputs(self._h, _p(nxt), _p(act))
        return rc, nxt, act

    def simulate_rollout(self, policy, n, steps, want_final=True):
        """ust_simulate_rollout on the resident snapshot. Returns (rc, steps_done, [counters-dict per step], final dict)."""
        hist = (abi.Counters * max(steps, 1))()
        fin = {"state": np.zeros(n, np.uint8), "flags": np.zeros(n, np.uint32), "pod_rev": np.zeros(n, np.int32)} if want_final else None
        done = C.c_int32(0)
        rc = self._lib.ust_simulate_rollout(
            self._h, C.addressof(policy) if policy is not None else None, int(steps), C.addressof(hist),
            _p(fin["state"]) if fin else None, _p(fin["flags"]) if fin else None, _p(fin["pod_rev"]) if fin else None,
            C.addressof(done))
        return rc, int(done.value), [hist[k].as_dict() for k in range(steps)], fin

    def simulate_rollout_timed(self, policy, options, n, steps, want_final=True):
        """ust_simulate_rollout_timed on the resident snapshot (options: abi.SimOptions)."""
        hist = (abi.Counters * max(steps, 1))()
        fin = {"state": np.zeros(n, np.uint8), "flags": np.zeros(n, np.uint32), "pod_rev": np.zeros(n, np.int32)} if want_final else None
        done = C.c_int32(0)
        rc = self._lib.ust_simulate_rollout_timed(
            self._h, C.addressof(policy) if policy is not None else None, C.addressof(options), int(steps), C.addressof(hist),
            _p(fin["state"]) if fin else None, _p(fin["flags"]) if fin else None, _p(fin["pod_rev"]) if fin else None,
            C.addressof(done))
        return rc, int(done.value), [hist[k].as_dict() for k in range(steps)], fin

    def apply_state_device(self, policy, n, state, flags, pod_rev, ds_idx, n_ds, ds_rev, next_state, actions,
                           outcome=None, pods=None, counters=None, stream=None):
        """Device-pointer entry point (ust_apply_state_device). Arguments are raw device addresses."""
        ps = None
        if pods is not None:
            ps = abi.Pods(int(pods[0]), int(pods[1]), int(pods[2]))
        rc = self._lib.ust_apply_state_device(
            self._h, C.addressof(policy) if policy is not None else None, int(n), _p(state), _p(flags), _p(pod_rev),
            _p(ds_idx), int(n_ds), _p(ds_rev), C.addressof(ps) if ps is not None else None, _p(next_state),
            _p(actions), _p(outcome), _p(counters), _p(stream))
        if rc:
            raise UstError(rc, self.last_error())

    def build_state(self, state, ds_idx, ds_desired):
        cnt = abi.Counters()
        rc = self._lib.ust_build_state(self._h, int(state.shape[0]), _p(state), _p(ds_idx), int(ds_desired.shape[0]),
                                       _p(ds_desired), C.addressof(cnt))
        return rc, cnt.as_dict()

    def build_state_uids(self, state, owner_uid, ds_uid, ds_desired):
        """BuildState with the owner join on the device: owner_uid (n, 2) uint64, ds_uid (n_ds, 2) uint64.
        Returns (rc, ds_idx per pod, counters-dict)."""
        n = int(state.shape[0])
        owner_uid = np.ascontiguousarray(owner_uid, dtype=np.uint64).reshape(n, 2)
        ds_uid = np.ascontiguousarray(ds_uid, dtype=np.uint64).reshape(-1, 2)
        ds_desired = np.ascontiguousarray(ds_desired, dtype=np.int32)
        ds_idx = np.full(n, -3, np.int32)
        cnt = abi.Counters()
        rc = self._lib.ust_build_state_uids(self._h, n, _p(state), _p(owner_uid), int(ds_uid.shape[0]), _p(ds_uid),
                                            _p(ds_desired), _p(ds_idx), C.addressof(cnt))
        return rc, ds_idx, cnt.as_dict()

    def comm_init(self, rank, world, unique_id_bytes):
        buf = (C.c_char * abi.UST_UNIQUE_ID_BYTES).from_buffer_copy(unique_id_bytes) if unique_id_bytes else None
        rc = self._lib.ust_comm_init(self._h, rank, world, buf)
        if rc:
            raise UstError(rc, self.last_error())

    def comm_set_mode(self, mode):
        rc = self._lib.ust_comm_set_mode(self._h, mode)
        if rc:
            raise UstError(rc, self.last_error())


def get_unique_id():
    buf = (C.c_char * abi.UST_UNIQUE_ID_BYTES)()
    rc = load().ust_get_unique_id(buf)
    if rc:
        raise UstError(rc, load().ust_create_error().decode())
    return bytes(buf)
